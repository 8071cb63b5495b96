// TEST INFRASTRUCTURE (oracle) -- force-included (-include) in front of every reference
// C++ translation unit that oracle/build_ref.sh compiles in place.  It only declares what
// the headers we skip (via their include guards, see build_ref.sh) would have declared.
#ifndef MB2_REF_PREINCLUDE_HPP
#define MB2_REF_PREINCLUDE_HPP
#include <string>
#include <cstdlib>
#include <sstream>
#include <fstream>
#include <iostream>
#include <opencv2/core/core.hpp>
struct DescriptorsParameters;     // descriptors_parameters.hpp is skipped (pulls every descriptor family)
struct DominantOrientationParams;
// TILDE/c++/src/libTILDE.hpp is skipped (pyramid.cpp:16 only needs this one symbol to compile).
template <class... A> inline cv::Mat getTILDEResponce(A&&...) { std::abort(); }
#endif
