// extractor.h -- the extractor handle (device workspace of one ORBextractor object)
#pragma once
#include <vector>

#include "common.cuh"

struct orbr_state;
void orbr_release(struct orbx_handle* h);   // replay.cu

struct orbx_handle {
    orbx_config cfg{};
    orb::ExtractGeom geom{};
    int cur_w = -1, cur_h = -1;
    int fast_variant = 1; // 1 (default): k_fast_cells_v2; 0: the round-1 k_fast_cells (ORB_FAST_VARIANT=0 at orbx_create)
    std::vector<uint2> cells_host;   // k_fast_cells_v2: one record per non-empty FAST cell
    uint2* d_cells = nullptr;
    size_t cells_slots = 0;
    int fast_R = 0, fast_PW = 0, fast_LW = 0;
    int blur_variant = 1;    // 1: k_blur<true> (IDP.4A horizontal pass, default), 0: k_blur<false>
    int resize_variant = 1;  // 1: k_resize_v2 (default), 0: k_resize
    bool fast_tma = false;   // ORB_FAST_TMA=1: k_fast_cells_v2<29, true> (window rows by cp.async.bulk + mbarrier)
    int qt_variant = 1;   // 1 (default): k_quadtree_v1; 0: k_quadtree (ORB_QT_VARIANT=0 at orbx_create)
    cudaStream_t stream = nullptr;
    static const int kProfRing = 32;
    cudaEvent_t evr[kProfRing][8] = {};   // ring of per-batch stage events
    cudaEvent_t* ev = evr[0];
    int prof_count = 0;
    bool profiling = false;
    // ORBextractor.h:97-108 tables
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> quota, umax;
    std::vector<int2> taps_host;
    // device workspace (sized for max_width x max_height x max_batch)
    uint8_t* d_pyr = nullptr;      // level-major pyramids (mvImagePyramid)
    uint8_t* d_blur = nullptr;     // blurred pyramids
    size_t pyr_bytes = 0;
    uint32_t* d_cand = nullptr;    // FAST candidates, packed x|y<<12|score<<24
    uint32_t* d_sort = nullptr;    // global fallback for the quadtree sort
    uint32_t* d_lvl_kp = nullptr;  // DistributeOctTree output per level
    int* d_slot = nullptr;         // output row of every emitted keypoint
    int* d_cand_cnt = nullptr;
    int* d_lvl_cnt = nullptr;
    int* d_nkp = nullptr;          // [max_batch] n, [max_batch] mono, [max_batch+1] offsets
    int* d_mono = nullptr;
    int* d_offsets = nullptr;
    int* d_err = nullptr;
    int2* d_taps = nullptr;
    orbx_keypoint* d_kps = nullptr;  // compact results of the last batch
    uint8_t* d_desc = nullptr;
    float* d_uright = nullptr;       // Frame::mvuRight / mvDepth of the last stereo batch (row-aligned with d_kps)
    float* d_depth = nullptr;
    int* d_sad = nullptr;
    uint32_t* d_pat = nullptr;       // rBRIEF pattern, 256 packed pairs transposed for k_orient_describe
    void* d_node_scratch = nullptr;
    uint8_t* d_stage = nullptr;
    uint8_t* d_po = nullptr;         // edge lists / results of orbo_pose_optimization_frames
    size_t node_scratch_bytes = 0, stage_bytes = 0, po_bytes = 0;
    bool stereo_valid = false;   // d_uright/d_depth belong to the last extracted batch
    size_t cand_slots = 0, kp_slots = 0, sort_slots = 0, taps_slots = 0, out_rows = 0;
    int* h_counts = nullptr;       // pinned
    bool counts_valid = false;
    int q_total_bound = 0, q_frame_bound = 0, rows_bound = 0;   // orbm_set_device_query_bounds: sync-free device-pointer searches
    bool capturing = false;   // between orbx_graph_begin and orbx_graph_end
    int batch_status = 0;   // 0 ok; 1 FAST candidate overflow, 2 quadtree node overflow of the last batch (truncated results are never served)
    int last_batch = 0;
    // quadtree launch plan: up to 4 level groups with their own shared-memory size, run on parallel
    // streams (forked from / joined into `stream`)
    struct QtGroup { int level_begin, level_end, sort_cap; size_t smem; };
    QtGroup qt_groups[4] = {};
    int qt_max_groups = 3;   // level groups (launches) of the quadtree stage: {16384+, 8192, 2048} sort capacities; ORB_QT_GROUPS=4 adds a 4096 group
                             // (measured in round 2: 16 KB less shared memory for levels 2-4, but 0.317 instead of 0.278 ms per batch -- more
                             // co-resident CTAs contend for the same SMs' barrier / shared-memory latency)
    int qt_ngroups = 0;
    cudaStream_t aux_stream[3] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {};
    int qt_node_cap = 0, qt_nodes_in_smem = 1;
    size_t qt_node_stride = 0, order_smem_bytes = 0;
    orbr_state* replay = nullptr;   // orbr_submit / orbr_collect state (replay.cu)
};
