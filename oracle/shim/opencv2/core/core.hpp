// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product path.
//
// Minimal stand-in for the OpenCV 2.4.9 C++ API surface that the reference's hot-path
// sources touch.  It exists only so that oracle/build_ref.sh can compile
// the reference's own .cpp files *where they lie* under /root/reference into
// oracle/_ref/libmods_ref.so.  Numerics (GaussianBlur / resize / invert) forward to the
// restatement in oracle/cvmath.h; everything the hot path never calls aborts loudly.
#ifndef MB2_ORACLE_SHIM_CORE_HPP
#define MB2_ORACLE_SHIM_CORE_HPP

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <algorithm>

#include "../../../cvmath.h"

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_8UC3 (CV_8U + 16)
#define CV_32FC1 CV_32F
#define CV_32FC3 (CV_32F + 16)
#define CV_64FC1 CV_64F
#define CV_PI 3.1415926535897932384626433832795
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif

namespace cv {

[[noreturn]] inline void shim_unsupported(const char* what) {
  std::fprintf(stderr, "opencv shim: '%s' is outside the hot path and not implemented\n", what);
  std::abort();
}

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Point { int x, y; Point() : x(0), y(0) {} Point(int a, int b) : x(a), y(b) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
template <class T> using Ptr = std::shared_ptr<T>;

enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { DECOMP_LU = 0, DECOMP_SVD = 1 };

class Mat {
 public:
  int rows, cols;
  unsigned char* data;
  size_t step;
  int flags_type;
  std::shared_ptr<unsigned char> owner;

  Mat() : rows(0), cols(0), data(nullptr), step(0), flags_type(CV_8U) {}
  Mat(int r, int c, int t) : Mat() { create(r, c, t); }
  Mat(Size s, int t) : Mat() { create(s.height, s.width, t); }
  Mat(int r, int c, int t, const Scalar& s) : Mat() { create(r, c, t); *this = s; }
  Mat(int r, int c, int t, void* d, size_t st = 0) : rows(r), cols(c), data((unsigned char*)d), flags_type(t) {
    step = st ? st : (size_t)c * elemSize();
  }
  template <class T> explicit Mat(const std::vector<T>&) : Mat() { shim_unsupported("Mat(vector) (drawing code)"); }
  Mat operator()(const Rect&) const { shim_unsupported("Mat(Rect) (drawing code)"); }
  double dot(const Mat&) const { shim_unsupported("Mat::dot"); }
  void create(int r, int c, int t) {
    rows = r; cols = c; flags_type = t;
    step = (size_t)c * elemSize();
    size_t n = step * (size_t)r;
    owner.reset((unsigned char*)std::malloc(n ? n : 1), std::free);
    data = owner.get();
  }
  int type() const { return flags_type; }
  int depth() const { return flags_type & 7; }
  int channels() const { return (flags_type >> 4) + 1; }
  size_t elemSize1() const { int d = depth(); return d == CV_8U ? 1 : d == CV_64F ? 8 : 4; }
  size_t elemSize() const { return elemSize1() * channels(); }
  size_t total() const { return (size_t)rows * cols; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return true; }
  Size size() const { return Size(cols, rows); }

  template <class T> T* ptr(int r = 0) { return (T*)(data + step * r); }
  template <class T> const T* ptr(int r = 0) const { return (const T*)(data + step * r); }
  template <class T> T& at(int r, int c) { return ((T*)(data + step * r))[c]; }
  template <class T> const T& at(int r, int c) const { return ((const T*)(data + step * r))[c]; }
  template <class T> T& at(int i) { return ((T*)data)[i]; }
  template <class T> const T& at(int i) const { return ((const T*)data)[i]; }

  Mat clone() const {
    Mat m(rows, cols, flags_type);
    for (int r = 0; r < rows; r++) std::memcpy(m.data + m.step * r, data + step * r, (size_t)cols * elemSize());
    return m;
  }
  void copyTo(Mat& m) const { m = clone(); }
  static Mat zeros(int r, int c, int t) { Mat m(r, c, t); std::memset(m.data, 0, m.step * r); return m; }
  Mat& operator=(const Scalar& s) {
    size_t n = total() * channels();
    switch (depth()) {
      case CV_8U: for (size_t i = 0; i < n; i++) data[i] = (unsigned char)s.val[0]; break;
      case CV_32F: for (size_t i = 0; i < n; i++) ((float*)data)[i] = (float)s.val[0]; break;
      case CV_64F: for (size_t i = 0; i < n; i++) ((double*)data)[i] = s.val[0]; break;
      default: shim_unsupported("Mat=Scalar depth");
    }
    return *this;
  }
  Mat mul(const Mat& o) const {
    Mat m(rows, cols, flags_type);
    if (depth() != CV_32F) shim_unsupported("Mat::mul depth");
    size_t n = total();
    for (size_t i = 0; i < n; i++) ((float*)m.data)[i] = ((const float*)data)[i] * ((const float*)o.data)[i];
    return m;
  }
  Mat t() const {
    Mat m(cols, rows, flags_type);
    if (depth() != CV_64F) shim_unsupported("Mat::t depth");
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) m.at<double>(c, r) = at<double>(r, c);
    return m;
  }
  void convertTo(Mat& m, int t) const {
    if ((t & 7) == depth()) { m = clone(); return; }
    shim_unsupported("Mat::convertTo");
  }
};

template <class T> struct MatCommaInit_;
template <class T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, sizeof(T) == 8 ? CV_64F : CV_32F) {}
  Mat_(const Mat& m) : Mat(m) {}
};
template <class T> struct MatCommaInit_ {
  Mat_<T> m; int i;
  MatCommaInit_(const Mat_<T>& mm, T v) : m(mm), i(0) { m.template at<T>(i++) = v; }
  template <class U> MatCommaInit_& operator,(U v) { m.template at<T>(i++) = (T)v; return *this; }
  operator Mat_<T>() const { return m; }
  operator Mat() const { return m; }
};
template <class T, class U> inline MatCommaInit_<T> operator<<(const Mat_<T>& m, U v) { return MatCommaInit_<T>(m, (T)v); }

inline Mat mat_binop(const Mat& a, const Mat& b, int op) {
  if (a.depth() != CV_32F || b.depth() != CV_32F) shim_unsupported("Mat +/- depth");
  Mat m(a.rows, a.cols, a.flags_type);
  size_t n = a.total() * a.channels();
  const float* x = (const float*)a.data; const float* y = (const float*)b.data; float* z = (float*)m.data;
  for (size_t i = 0; i < n; i++) z[i] = op ? x[i] - y[i] : x[i] + y[i];
  return m;
}
inline Mat operator+(const Mat& a, const Mat& b) { return mat_binop(a, b, 0); }
inline Mat operator-(const Mat& a, const Mat& b) { return mat_binop(a, b, 1); }
inline Mat operator*(double s, const Mat& a) {
  if (a.depth() != CV_32F) shim_unsupported("scalar*Mat depth");
  Mat m(a.rows, a.cols, a.flags_type);
  size_t n = a.total() * a.channels();
  for (size_t i = 0; i < n; i++) ((float*)m.data)[i] = (float)(((const float*)a.data)[i] * s);
  return m;
}
inline Mat operator*(const Mat& a, double s) { return s * a; }
inline Mat operator/(const Mat& a, double s) { return (1.0 / s) * a; }
inline Mat operator*(const Mat&, const Mat&) { shim_unsupported("Mat*Mat"); }
inline Mat operator+(const Mat&, double) { shim_unsupported("Mat+scalar (dead WLD branch, extrema.cpp:125,317)"); }

struct SVD {
  static void compute(const Mat&, Mat&, Mat&, Mat&, int = 0) { shim_unsupported("SVD::compute"); }
};
inline void SVDecomp(const Mat&, Mat&, Mat&, Mat&, int = 0) { shim_unsupported("SVDecomp"); }

inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0,
                         int borderType = BORDER_DEFAULT) {
  if (src.depth() != CV_32F || src.channels() != 1) shim_unsupported("GaussianBlur type");
  if (borderType != BORDER_REPLICATE && borderType != BORDER_REFLECT_101) shim_unsupported("GaussianBlur border");
  if (sigmaY <= 0) sigmaY = sigmaX;
  Mat out = (dst.data == src.data) ? src : Mat(src.rows, src.cols, src.type());
  std::vector<float> kx = cvmath::gauss_kernel(ksize.width, sigmaX);
  std::vector<float> ky = cvmath::gauss_kernel(ksize.height, sigmaY);
  if (borderType == BORDER_REPLICATE) cvmath::sep_filter((const float*)src.data, (float*)out.data, src.rows, src.cols, kx, ky);
  else cvmath::sep_filter_reflect101((const float*)src.data, (float*)out.data, src.rows, src.cols, kx, ky);
  dst = out;
}

inline void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interp = INTER_LINEAR) {
  if (src.depth() != CV_32F || src.channels() != 1 || dsize.width != 0 || fx != 0.5 || fy != 0.5 || interp != INTER_LINEAR)
    shim_unsupported("resize other than x0.5 INTER_LINEAR");
  int orows, ocols;
  cvmath::half_size(src.rows, src.cols, &orows, &ocols);
  Mat out(orows, ocols, src.type());
  cvmath::resize_half((const float*)src.data, src.rows, src.cols, (float*)out.data);
  dst = out;
}

inline double invert(const Mat& src, Mat& dst, int = DECOMP_LU) {
  if (src.depth() != CV_64F || src.rows != 3 || src.cols != 3) shim_unsupported("invert other than 3x3 double");
  Mat out(3, 3, CV_64F);
  bool ok = cvmath::invert3x3((const double*)src.data, (double*)out.data);
  dst = out;
  return ok ? 1.0 : 0.0;
}

inline void warpAffine(const Mat& src, Mat& dst, const Mat& M, Size dsize, int flags = INTER_LINEAR, int border = BORDER_CONSTANT,
                       const Scalar& bv = Scalar()) {
  if (src.depth() != CV_32F || src.channels() != 1 || flags != INTER_LINEAR || border != BORDER_CONSTANT || M.depth() != CV_64F)
    shim_unsupported("warpAffine other than CV_32FC1 / INTER_LINEAR / BORDER_CONSTANT");
  Mat out(dsize.height, dsize.width, src.type());
  cvmath::warp_affine_linear((const float*)src.data, src.rows, src.cols, (const double*)M.data, (float*)out.data, dsize.height, dsize.width, (float)bv.val[0]);
  dst = out;
}
inline void warpPerspective(const Mat&, Mat&, const Mat&, Size, int = INTER_LINEAR, int = BORDER_CONSTANT, const Scalar& = Scalar()) {
  shim_unsupported("warpPerspective");
}
inline void split(const Mat&, std::vector<Mat>&) { shim_unsupported("split"); }
inline void gemm(const Mat&, const Mat&, double, const Mat&, double, Mat&, int = 0) { shim_unsupported("gemm"); }
// drawing / colour API referenced by matching.cpp's Draw* functions (outside the hot path, never called by the oracle)
#define CV_AA 16
#define CV_GRAY2RGB 8
#define CV_GRAY2BGR 8
inline Scalar operator*(const Scalar& s, double k) { return Scalar(s.val[0] * k, s.val[1] * k, s.val[2] * k, s.val[3] * k); }
inline Scalar operator+(const Scalar& a, const Scalar& b) { return Scalar(a.val[0] + b.val[0], a.val[1] + b.val[1], a.val[2] + b.val[2], a.val[3] + b.val[3]); }
template <class... A> inline void circle(A&&...) { shim_unsupported("circle"); }
template <class... A> inline void line(A&&...) { shim_unsupported("line"); }
template <class... A> inline void ellipse(A&&...) { shim_unsupported("ellipse"); }
template <class... A> inline void polylines(A&&...) { shim_unsupported("polylines"); }
template <class... A> inline void cvtColor(A&&...) { shim_unsupported("cvtColor"); }
template <class... A> inline bool clipLine(A&&...) { shim_unsupported("clipLine"); }
template <class... A> inline void addWeighted(A&&...) { shim_unsupported("addWeighted"); }
template <class... A> inline void rectangle(A&&...) { shim_unsupported("rectangle"); }
template <class... A> inline void putText(A&&...) { shim_unsupported("putText"); }
inline bool imwrite(const std::string&, const Mat&) { shim_unsupported("imwrite"); }
inline Mat imread(const std::string&, int = 1) { shim_unsupported("imread"); }

}  // namespace cv
#endif
