// ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the bag-of-words row:
//   DBoW2::TemplatedVocabulary::transform  Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1131-1194, 1218-1260
//   DBoW2::BowVector::addWeight / normalize Thirdparty/DBoW2/DBoW2/BowVector.cpp:32-83
//   DBoW2::FeatureVector::addFeature        Thirdparty/DBoW2/DBoW2/FeatureVector.cpp
//   FORB::distance                          Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101
//   ORBmatcher::SearchByBoW(KeyFrame*,Frame&,...)  src/ORBmatcher.cc:161-290, ComputeThreeMaxima :1749-1790
// Parity: PINNED.  The reference ships no tests or vectors for this row, so oracle/_ref compiles DBoW2 itself (TemplatedVocabulary<FORB>,
// BowVector, FeatureVector, FORB, ScoringObject -- vendored under Thirdparty/DBoW2) and both SearchByBoW overloads; a synthetic tree is
// written in the ORBvoc.txt format, loaded with the reference's loadFromTextFile and transformed there: BowVector doubles, FeatureVector
// and match arrays must equal this file's bit for bit (tests/test_oracle_vs_ref.py).  std::map semantics included (words / nodes
// ascending, double sums in feature order).
#include "oracle.h"
#include "cvprim.hpp"
#include <map>
#include <vector>
#include <cmath>

using orc::hamming256;

struct orc_vocab {
    int k, L, n;
    std::vector<uint8_t> desc; std::vector<int> cb, cc, ch, wid; std::vector<double> w; std::vector<int> parent;
};

extern "C" orc_vocab* orc_vocab_create(const olf_vocab_desc* v) {
    if (!v || v->n_nodes < 1) return nullptr;
    orc_vocab* o = new orc_vocab();
    o->k = v->k; o->L = v->L; o->n = v->n_nodes;
    o->desc.assign(v->node_desc, v->node_desc + (size_t)v->n_nodes * 32);
    o->cb.assign(v->child_begin, v->child_begin + v->n_nodes); o->cc.assign(v->child_count, v->child_count + v->n_nodes);
    int nch = 0; for (int i = 0; i < v->n_nodes; ++i) nch = std::max(nch, v->child_begin[i] + v->child_count[i]);
    o->ch.assign(v->children, v->children + nch);
    o->wid.assign(v->word_id, v->word_id + v->n_nodes); o->w.assign(v->weight, v->weight + v->n_nodes);
    return o;
}
extern "C" void orc_vocab_destroy(orc_vocab* v) { delete v; }

// transform(feature, word_id, weight, &nid, levelsup)  :1218-1260
extern "C" int orc_bow_transform(orc_vocab* V, const uint8_t* desc, int n, int levelsup, int* word_id, double* weight, int* node_id) {
    if (!V || n < 0) return OLF_ERR_ARG;
    const int nid_level = V->L - levelsup;
    for (int f = 0; f < n; ++f) {
        const uint8_t* feature = desc + (size_t)f * 32;
        int nid = 0;                                            // if(nid_level <= 0 && nid != NULL) *nid = 0; // root
        int final_id = 0, current_level = 0;
        do {
            ++current_level;
            const int* nodes = &V->ch[V->cb[final_id]];
            const int nn = V->cc[final_id];
            final_id = nodes[0];
            double best_d = hamming256(feature, &V->desc[(size_t)final_id * 32]);
            for (int c = 1; c < nn; ++c) {
                const int id = nodes[c];
                const double d = hamming256(feature, &V->desc[(size_t)id * 32]);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (V->cc[final_id] != 0);                         // !isLeaf()
        word_id[f] = V->wid[final_id]; weight[f] = V->w[final_id]; node_id[f] = nid;
    }
    return OLF_OK;
}

// transform(features, v, fv, levelsup) with TF_IDF weighting and L1_NORM scoring (mustNormalize) :1131-1194
extern "C" int orc_bow_assemble(const int* word_id, const double* weight, const int* node_id, int n,
                                int* bow_word, double* bow_value, int* n_words, int* fv_node, int* fv_begin, int* fv_index, int* n_nodes) {
    std::map<unsigned, double> v;                               // BowVector
    std::map<unsigned, std::vector<unsigned>> fv;               // FeatureVector
    for (int i = 0; i < n; ++i) {
        if (weight[i] > 0) {                                    // not stopped
            auto vit = v.lower_bound((unsigned)word_id[i]);     // addWeight
            if (vit != v.end() && !(v.key_comp()((unsigned)word_id[i], vit->first))) vit->second += weight[i];
            else v.insert(vit, std::make_pair((unsigned)word_id[i], weight[i]));
            fv[(unsigned)node_id[i]].push_back((unsigned)i);    // addFeature
        }
    }
    double norm = 0.0;                                          // normalize(L1)
    for (auto& e : v) norm += std::fabs(e.second);
    if (norm > 0.0) for (auto& e : v) e.second /= norm;
    int k = 0;
    for (auto& e : v) { bow_word[k] = (int)e.first; bow_value[k] = e.second; ++k; }
    *n_words = k;
    int m = 0, pos = 0;
    for (auto& e : fv) { fv_node[m] = (int)e.first; fv_begin[m] = pos; for (unsigned idx : e.second) fv_index[pos++] = (int)idx; ++m; }
    fv_begin[m] = pos;
    *n_nodes = m;
    return OLF_OK;
}

static void three_maxima_bow(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {   // src/ORBmatcher.cc:1749-1790
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
}

// SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches)  src/ORBmatcher.cc:161-290
extern "C" int orc_search_by_bow(const olf_bow_match_args* a, int* match_f, int* nmatches_out) {
    if (!a || !match_f || !nmatches_out) return OLF_ERR_ARG;
    for (int i = 0; i < a->n_f; ++i) match_f[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[OLF_HISTO_LENGTH];
    const float factor = 1.0f / OLF_HISTO_LENGTH;
    int ik = 0, jf = 0;
    while (ik < a->kf_n_nodes && jf < a->f_n_nodes) {
        if (a->kf_fv_node[ik] == a->f_fv_node[jf]) {
            for (int p = a->kf_fv_begin[ik]; p < a->kf_fv_begin[ik + 1]; ++p) {
                const int realIdxKF = a->kf_fv_index[p];
                if (!a->kf_has_point[realIdxKF]) continue;
                const uint8_t* dKF = a->kf_desc + (size_t)realIdxKF * 32;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int q = a->f_fv_begin[jf]; q < a->f_fv_begin[jf + 1]; ++q) {
                    const int realIdxF = a->f_fv_index[q];
                    if (match_f[realIdxF] >= 0) continue;
                    const int dist = hamming256(dKF, a->f_desc + (size_t)realIdxF * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= OLF_TH_LOW) {
                    if ((float)bestDist1 < a->nn_ratio * (float)bestDist2) {
                        match_f[bestIdxF] = realIdxKF;
                        if (a->check_orientation) {
                            float rot = a->kf_kps_un[realIdxKF].angle - a->f_kps[bestIdxF].angle;
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)roundf(rot * factor);
                            if (bin == OLF_HISTO_LENGTH) bin = 0;
                            rotHist[bin].push_back(bestIdxF);
                        }
                        nmatches++;
                    }
                }
            }
            ++ik; ++jf;
        } else if (a->kf_fv_node[ik] < a->f_fv_node[jf]) {
            while (ik < a->kf_n_nodes && a->kf_fv_node[ik] < a->f_fv_node[jf]) ++ik;        // lower_bound
        } else {
            while (jf < a->f_n_nodes && a->f_fv_node[jf] < a->kf_fv_node[ik]) ++jf;
        }
    }
    if (a->check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima_bow(rotHist, OLF_HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < OLF_HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { match_f[idx] = -1; nmatches--; }
        }
    }
    *nmatches_out = nmatches;
    return OLF_OK;
}

// SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12)  src/ORBmatcher.cc:524-657 (kf = pKF1, f = pKF2)
extern "C" int orc_search_by_bow_kf(const olf_bow_match_args* a, const uint8_t* has_point2, int* matches12, int* nmatches_out) {
    if (!a || !has_point2 || !matches12 || !nmatches_out) return OLF_ERR_ARG;
    for (int i = 0; i < a->n_kf; ++i) matches12[i] = -1;
    std::vector<uint8_t> vbMatched2(std::max(a->n_f, 1), 0);
    int nmatches = 0;
    std::vector<int> rotHist[OLF_HISTO_LENGTH];
    const float factor = 1.0f / OLF_HISTO_LENGTH;
    int i1n = 0, i2n = 0;
    while (i1n < a->kf_n_nodes && i2n < a->f_n_nodes) {
        if (a->kf_fv_node[i1n] == a->f_fv_node[i2n]) {
            for (int p = a->kf_fv_begin[i1n]; p < a->kf_fv_begin[i1n + 1]; ++p) {
                const int idx1 = a->kf_fv_index[p];
                if (!a->kf_has_point[idx1]) continue;
                const uint8_t* d1 = a->kf_desc + (size_t)idx1 * 32;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int q = a->f_fv_begin[i2n]; q < a->f_fv_begin[i2n + 1]; ++q) {
                    const int idx2 = a->f_fv_index[q];
                    if (vbMatched2[idx2] || !has_point2[idx2]) continue;
                    const int dist = hamming256(d1, a->f_desc + (size_t)idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < OLF_TH_LOW) {
                    if ((float)bestDist1 < a->nn_ratio * (float)bestDist2) {
                        matches12[idx1] = bestIdx2;
                        vbMatched2[bestIdx2] = 1;
                        if (a->check_orientation) {
                            float rot = a->kf_kps_un[idx1].angle - a->f_kps[bestIdx2].angle;
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)roundf(rot * factor);
                            if (bin == OLF_HISTO_LENGTH) bin = 0;
                            rotHist[bin].push_back(idx1);
                        }
                        nmatches++;
                    }
                }
            }
            ++i1n; ++i2n;
        } else if (a->kf_fv_node[i1n] < a->f_fv_node[i2n]) {
            while (i1n < a->kf_n_nodes && a->kf_fv_node[i1n] < a->f_fv_node[i2n]) ++i1n;
        } else {
            while (i2n < a->f_n_nodes && a->f_fv_node[i2n] < a->kf_fv_node[i1n]) ++i2n;
        }
    }
    if (a->check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima_bow(rotHist, OLF_HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < OLF_HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) { matches12[idx1] = -1; nmatches--; }
        }
    }
    *nmatches_out = nmatches;
    return OLF_OK;
}
