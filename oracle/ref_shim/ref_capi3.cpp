// oracle/ref_shim/ref_capi3.cpp -- TEST INFRASTRUCTURE.
//
// C entry points around the REFERENCE's own DBoW2 (Thirdparty/DBoW2, the pinned fork ORB-SLAM3 ships): oracle/Makefile (target ref3)
// compiles DBoW2/{BowVector, FeatureVector, ScoringObject, FORB}.cpp and DUtils/{Random, Timestamp}.cpp where they lie under
// /root/reference, unmodified, and this file instantiates TemplatedVocabulary<FORB::TDescriptor, FORB> (= ORBVocabulary,
// include/ORBVocabulary.h:30) from the reference's header -- over the miniature cv:: of ref_shim/opencv2 (cv::Mat as a u8 buffer;
// cv::FileStorage only parses) and two empty boost/serialization headers (ref_shim/dbow).  tests/test_oracle_vs_ref_dbow.py compares
// the oracle's orc_bow_transform (the walk down the tree, TF-IDF weights, L1 normalisation, BowVector / FeatureVector order) with
// what this library computes on the reference's own Vocabulary/ORBvoc.txt.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

namespace {
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;   // include/ORBVocabulary.h:30

struct OpenVoc : public ORBVocabulary {
    int nodes() const { return (int)m_nodes.size(); }
    void one(const cv::Mat& f, DBoW2::WordId& id, DBoW2::WordValue& w, DBoW2::NodeId& nid, int levelsup) const { transform(f, id, w, &nid, levelsup); }
};

cv::Mat row_mat(const uint8_t* d) {
    cv::Mat m(1, 32, CV_8U);
    std::memcpy(m.data, d, 32);
    return m;
}
}  // namespace

extern "C" {

void* ref3_voc_load_text(const char* path) {
    OpenVoc* v = new OpenVoc();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }   // System.cc:120
    return v;
}
void ref3_voc_destroy(void* v) { delete (OpenVoc*)v; }

void ref3_voc_info(void* h, int* out6) {
    OpenVoc* v = (OpenVoc*)h;
    out6[0] = v->getBranchingFactor(); out6[1] = v->getDepthLevels(); out6[2] = (int)v->getScoringType(); out6[3] = (int)v->getWeightingType();
    out6[4] = (int)v->size(); out6[5] = v->nodes();
}

// Frame::ComputeBoW (Frame.cc:984-997): vCurrentDesc = Converter::toDescriptorVector(mDescriptors) (one 1 x 32 Mat per row);
// mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4).  bow_* = mBowVec in map order; feat_node[i] = the mFeatVec key that
// lists feature i (-1: none); ascending[0] = 1 when every mFeatVec list is in ascending feature order.  Returns mBowVec.size().
int ref3_transform(void* h, const uint8_t* desc, int n, int levelsup, int* bow_word, double* bow_weight, int* feat_node, int* ascending) {
    OpenVoc* v = (OpenVoc*)h;
    std::vector<cv::Mat> feats;
    feats.reserve(n);
    for (int i = 0; i < n; ++i) feats.push_back(row_mat(desc + 32 * (size_t)i));
    DBoW2::BowVector bow;
    DBoW2::FeatureVector fv;
    v->transform(feats, bow, fv, levelsup);
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = bow.begin(); it != bow.end(); ++it, ++k) { bow_word[k] = (int)it->first; bow_weight[k] = it->second; }
    for (int i = 0; i < n; ++i) feat_node[i] = -1;
    *ascending = 1;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        for (size_t j = 0; j < it->second.size(); ++j) {
            feat_node[it->second[j]] = (int)it->first;
            if (j > 0 && it->second[j] <= it->second[j - 1]) *ascending = 0;
        }
    return k;
}

// one feature: TemplatedVocabulary::transform(feature, id, weight, &nid, levelsup) (:1216-1258), the walk behind the call above; also
// checked against the public transform(const TDescriptor&) / getWordWeight
int ref3_transform_one(void* h, const uint8_t* d32, int levelsup, int* word, double* weight, int* node) {
    OpenVoc* v = (OpenVoc*)h;
    DBoW2::WordId id = 0;
    DBoW2::WordValue w = 0;
    DBoW2::NodeId nid = 0;
    v->one(row_mat(d32), id, w, nid, levelsup);
    *word = (int)id; *weight = w; *node = (int)nid;
    return v->transform(row_mat(d32)) == id && v->getWordWeight(id) == w;
}

// ORBVocabulary::score (KeyFrameDatabase.cc:150, 272 ...): the vocabulary's scoring object (L1_NORM for ORBvoc.txt) on two BowVectors
double ref3_score(void* h, int n1, const int* w1, const double* v1, int n2, const int* w2, const double* v2) {
    OpenVoc* v = (OpenVoc*)h;
    DBoW2::BowVector a, b;
    for (int i = 0; i < n1; ++i) a.insert(a.end(), std::make_pair((DBoW2::WordId)w1[i], v1[i]));
    for (int i = 0; i < n2; ++i) b.insert(b.end(), std::make_pair((DBoW2::WordId)w2[i], v2[i]));
    return v->score(a, b);
}

int ref3_forb_distance(const uint8_t* a, const uint8_t* b) { return DBoW2::FORB::distance(row_mat(a), row_mat(b)); }

}  // extern "C"
