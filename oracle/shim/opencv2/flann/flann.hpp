// Shim header (TEST INFRASTRUCTURE): stand-in for cv::flann::Index so that the reference's matching/matching.cpp compiles in place.
// OpenCV's FLANN is not in /root/reference and cannot be built here.  Every index type is answered by an EXACT
// linear k-NN on squared L2 (what `vector_matcher=linear` asks FLANN for; BASELINE config 2 says "FLANN-off brute NN"); equal
// distances are ordered by the lower train index (FLANN's tie order is unspecified).  Descriptor entries are integers
// 0..255 held in floats, so the float sum cvflann::L2<float> forms is an exact integer < 2^24 in any order: it is evaluated in
// 32-bit integer arithmetic (vectorisable) when that holds, in float otherwise.  Queries run on all OpenMP threads.
// Binary descriptors (u8 rows, FLANN_DIST_HAMMING; MatchFLANNDistance): exact linear k-NN on the bit count, int distances.
#ifndef MB2_ORACLE_SHIM_FLANN_HPP
#define MB2_ORACLE_SHIM_FLANN_HPP
#include "../core/core.hpp"
#include <utility>
namespace cvflann {
enum flann_algorithm_t { FLANN_INDEX_LINEAR = 0, FLANN_INDEX_KDTREE = 1, FLANN_INDEX_KMEANS = 2, FLANN_INDEX_COMPOSITE = 3, FLANN_INDEX_KDTREE_SINGLE = 4,
                         FLANN_INDEX_HIERARCHICAL = 5, FLANN_INDEX_LSH = 6, FLANN_INDEX_SAVED = 254, FLANN_INDEX_AUTOTUNED = 255 };
enum flann_distance_t { FLANN_DIST_EUCLIDEAN = 1, FLANN_DIST_L2 = 1, FLANN_DIST_MANHATTAN = 2, FLANN_DIST_L1 = 2, FLANN_DIST_HAMMING = 9 };
}
namespace cv { namespace flann {
struct IndexParams { template <class... A> IndexParams(A&&...) {} };
typedef IndexParams KDTreeIndexParams, CompositeIndexParams, AutotunedIndexParams, KMeansIndexParams, LshIndexParams, LinearIndexParams,
    HierarchicalClusteringIndexParams;
struct SearchParams { template <class... A> SearchParams(A&&...) {} };
class Index {
 public:
  Mat feats;
  Index() {}
  bool hamming = false;
  Index(const Mat& features, const IndexParams&, cvflann::flann_distance_t d = cvflann::FLANN_DIST_L2) : feats(features) {
    hamming = d == cvflann::FLANN_DIST_HAMMING && features.depth() == CV_8U;   // binary descriptors (MatchFLANNDistance, matching.cpp:607)
    if (!hamming && (d != cvflann::FLANN_DIST_L2 || features.depth() != CV_32F)) shim_unsupported("flann::Index other than float L2 / u8 Hamming");
  }
  void release() { feats = Mat(); }
  void knnSearch(const Mat& queries, Mat& indices, Mat& dists, int knn, const SearchParams& = SearchParams()) {
    const int nq = queries.rows, nt = feats.rows, D = feats.cols;
    if (nt < knn) shim_unsupported("flann knnSearch with fewer trains than neighbours (undefined in the reference)");
    if (hamming) {   // cvflann::Hamming: number of differing bits, int distances (CV_32S); exact linear scan, lower train index first
      if (queries.depth() != CV_8U || queries.cols != D) shim_unsupported("Hamming knnSearch query type");
      indices = Mat(nq, knn, CV_32S); dists = Mat(nq, knn, CV_32S);
#pragma omp parallel
      {
        std::vector<std::pair<int, int> > d(nt);
#pragma omp for schedule(dynamic, 16)
        for (int i = 0; i < nq; i++) {
          const unsigned char* a = queries.ptr<unsigned char>(i);
          for (int j = 0; j < nt; j++) {
            const unsigned char* b = feats.ptr<unsigned char>(j);
            int s = 0;
            for (int e = 0; e < D; e++) s += __builtin_popcount((unsigned)(a[e] ^ b[e]));
            d[j] = std::make_pair(s, j);
          }
          std::partial_sort(d.begin(), d.begin() + knn, d.end());
          int* ir = indices.ptr<int>(i); int* dr = dists.ptr<int>(i);
          for (int j = 0; j < knn; j++) { ir[j] = d[j].second; dr[j] = d[j].first; }
        }
      }
      return;
    }
    indices = Mat(nq, knn, CV_32S); dists = Mat(nq, knn, CV_32F);
    bool integral = true;
    for (int i = 0; i < nq && integral; i++) { const float* r = queries.ptr<float>(i); for (int e = 0; e < D; e++) if (!(r[e] >= 0 && r[e] <= 255 && r[e] == (float)(int)r[e])) { integral = false; break; } }
    for (int i = 0; i < nt && integral; i++) { const float* r = feats.ptr<float>(i); for (int e = 0; e < D; e++) if (!(r[e] >= 0 && r[e] <= 255 && r[e] == (float)(int)r[e])) { integral = false; break; } }
    std::vector<unsigned char> q8, t8;
    if (integral) {
      q8.resize((size_t)nq * D); t8.resize((size_t)nt * D);
      for (int i = 0; i < nq; i++) { const float* r = queries.ptr<float>(i); for (int e = 0; e < D; e++) q8[(size_t)i * D + e] = (unsigned char)r[e]; }
      for (int i = 0; i < nt; i++) { const float* r = feats.ptr<float>(i); for (int e = 0; e < D; e++) t8[(size_t)i * D + e] = (unsigned char)r[e]; }
    }
#pragma omp parallel
    {
      std::vector<std::pair<float, int> > d(nt);
#pragma omp for schedule(dynamic, 16)
      for (int i = 0; i < nq; i++) {
        if (integral) {
          const unsigned char* a = q8.data() + (size_t)i * D;
          for (int j = 0; j < nt; j++) {
            const unsigned char* b = t8.data() + (size_t)j * D;
            int s = 0;
            for (int e = 0; e < D; e++) { const int df = (int)a[e] - (int)b[e]; s += df * df; }
            d[j] = std::make_pair((float)s, j);
          }
        } else {
          const float* a = queries.ptr<float>(i);
          for (int j = 0; j < nt; j++) {
            const float* b = feats.ptr<float>(j);
            float s = 0;
            for (int e = 0; e < D; e++) { const float df = a[e] - b[e]; s += df * df; }
            d[j] = std::make_pair(s, j);
          }
        }
        std::partial_sort(d.begin(), d.begin() + knn, d.end());
        int* ir = indices.ptr<int>(i); float* dr = dists.ptr<float>(i);
        for (int j = 0; j < knn; j++) { ir[j] = d[j].second; dr[j] = d[j].first; }
      }
    }
  }
};
} }
#endif
