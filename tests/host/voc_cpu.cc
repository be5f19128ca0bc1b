// tests/host/voc_cpu.cc -- TEST INFRASTRUCTURE (CPU tier): Frame::ComputeBoW (host/Frame_bow_b200.cc, orbv_* answered by the oracle: voc_stub.cc)
// over a mock vocabulary object built from flat arrays and a mock frame; tests/test_host_voc_cpu.py compares the containers it fills.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "ref_skeleton_impl.h"

using namespace ORB_SLAM3;
extern "C" void voc_stub_set_frame(const uint8_t* desc, int n);
extern "C" int voc_stub_creates();

static std::string g_dir;
template <typename T>
static std::vector<T> rd(const std::string& name) {
    std::ifstream f(g_dir + "/" + name, std::ios::binary);
    if (!f) { std::fprintf(stderr, "missing input %s\n", name.c_str()); std::exit(2); }
    f.seekg(0, std::ios::end);
    const size_t n = (size_t)f.tellg();
    f.seekg(0);
    std::vector<T> v(n / sizeof(T));
    f.read((char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
    return v;
}
template <typename T>
static void wr(const std::string& name, const std::vector<T>& v) {
    std::ofstream f(g_dir + "/" + name, std::ios::binary);
    f.write((const char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: voc_cpu <dir>\n"); return 2; }
    g_dir = argv[1];
    auto co = rd<int>("child_offset.i32"), ci = rd<int>("child_ids.i32"), nw = rd<int>("node_word.i32"), meta = rd<int>("meta.i32");   // meta: L
    auto nd = rd<uint8_t>("node_desc.u8"), desc = rd<uint8_t>("desc.u8");
    auto wt = rd<double>("node_weight.f64");
    const int n = (int)nw.size(), N = (int)(desc.size() / 32);
    std::vector<ORBVocabulary::Node> nodes(n);
    for (int i = 0; i < n; ++i) {
        nodes[i].id = (DBoW2::NodeId)i;
        for (int c = co[i]; c < co[i + 1]; ++c) { nodes[i].children.push_back((DBoW2::NodeId)ci[c]); nodes[ci[c]].parent = (DBoW2::NodeId)i; }
        if (i > 0) { nodes[i].descriptor.create(1, 32, CV_8UC1); std::memcpy(nodes[i].descriptor.ptr(0), &nd[(size_t)i * 32], 32); }
        if (co[i] == co[i + 1]) { nodes[i].word_id = (DBoW2::WordId)nw[i]; nodes[i].weight = wt[i]; }
    }
    ORBVocabulary voc;
    voc.mock_set(meta[0], nodes);
    std::vector<int> out_i;
    std::vector<double> out_w;
    for (int pass = 0; pass < 2; ++pass) {      // two frames against the same vocabulary object: it is flattened once
        Frame F;
        F.N = pass == 0 ? N : N / 2;
        int dummy_extractor = 0;
        F.mpORBextractorLeft = reinterpret_cast<ORBextractor*>(&dummy_extractor);
        F.mpORBvocabulary = &voc;
        voc_stub_set_frame(desc.data(), F.N);
        F.ComputeBoW();
        const size_t before = F.mBowVec.size();
        F.ComputeBoW();                          // a second call is a no-op (mBowVec is not empty)
        if (F.mBowVec.size() != before) return 3;
        out_i.push_back((int)F.mBowVec.size());
        for (const auto& kv : F.mBowVec) { out_i.push_back((int)kv.first); out_w.push_back(kv.second); }
        out_i.push_back((int)F.mFeatVec.size());
        for (const auto& kv : F.mFeatVec) {
            out_i.push_back((int)kv.first); out_i.push_back((int)kv.second.size());
            for (unsigned int f : kv.second) out_i.push_back((int)f);
        }
    }
    out_i.push_back(voc_stub_creates());
    wr("out.i32", out_i); wr("out.f64", out_w);
    std::printf("voc_cpu ok\n");
    return 0;
}
