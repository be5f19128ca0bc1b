// tests/host/voc_access_probe.cc -- TEST INFRASTRUCTURE: the protected-member access host/Frame_bow_b200.cc uses to flatten the vocabulary, compiled against
// the reference's REAL Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h (tests/test_host_shim.py): it needs no change to that header.
#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;
struct VocabularyAccess : ORBVocabulary {
    typedef ORBVocabulary Base;
    static const std::vector<Node>& nodes(const Base& v) { return v.*(&VocabularyAccess::m_nodes); }
    static int depth_levels(const Base& v) { return v.*(&VocabularyAccess::m_L); }
};
int probe(const ORBVocabulary& v) {
    const auto& n = VocabularyAccess::nodes(v);
    int s = VocabularyAccess::depth_levels(v);
    for (const auto& nd : n) s += (int)nd.children.size() + (nd.isLeaf() ? (int)nd.word_id : 0) + (nd.descriptor.empty() ? 0 : nd.descriptor.ptr<unsigned char>()[0]) + (int)nd.weight;
    return s;
}
