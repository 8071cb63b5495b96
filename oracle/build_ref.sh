#!/bin/bash
# TEST INFRASTRUCTURE (oracle).  Compiles the reference's own hot-path sources IN PLACE from
# /root/reference (never copied) into oracle/_ref/libmods_ref.so:
#   detectors/helpers.cpp, detectors/affinedetectors/{pyramid,affine,scale-space-detector}.cpp,
#   matching/siftdesc.cpp, matching/matching.cpp (FLANN answered by the shim's exact linear k-NN), synth-detection.cpp,
#   detectors/mser/{extrema,utls,LL}/*,
#   degensac/{DegUtils,exp_ranF,exp_ranH,Ftools,hash,Htools,ranF,ranH,rtools,utools,lapwrap}.c
#   (= CMakeLists.txt:84-97 minus the unused ranH2el.c), matutls/*.c (ccmath subset).
# OpenCV 2.4.9 is replaced by oracle/shim (cv::Mat + restated GaussianBlur/resize/invert);
# LAPACK dsyev_/dgesvd_ come from the OpenBLAS bundled with opencv-python-headless.
# Heavy unrelated headers are skipped through their include guards.
set -e
REF=${MB2_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF/degensac" ]; then echo "build_ref: $REF not present, keeping prebuilt $OUT"; exit 0; fi
mkdir -p $OUT/obj
SP=$(python -c "import site; print(site.getsitepackages()[0])")
BLASDIR=$SP/opencv_python_headless.libs
BLAS=$(ls $BLASDIR/libopenblas*.so | head -1)
CXXFLAGS="-std=c++11 -O3 -ftree-vectorize -funroll-loops -fPIC -w -fpermissive -ffp-contract=off -fopenmp -DA64 \
 -DDESCRIPTORS_PARAMETERS_HPP -DMROGH_HPP -DSURFLIB_H -D_LIBTILDE_HPP_ \
 -include $HERE/ref_preinclude.hpp -I$HERE/shim -I$REF -I$REF/detectors -I$REF/degensac -I$REF/vlfeat -I$REF/detectors/mser/utls -I$REF/detectors/mser/LL -I$REF/detectors/mser/extrema"
CFLAGS="-O3 -ftree-vectorize -funroll-loops -fPIC -w -fcommon -ffp-contract=off -include $HERE/ref_preinclude_c.h -I$REF -I$REF/degensac -I$REF/matutls"
OBJS=""
cc_one() { # $1 = compiler+flags, $2 = source
  local o=$OUT/obj/$(echo "$2" | tr '/' '_').o
  if [ ! -f "$o" ] || [ "$REF/$2" -nt "$o" ] || [ "$HERE/shim/opencv2/core/core.hpp" -nt "$o" ] || [ "$HERE/cvmath.h" -nt "$o" ] || [ "$HERE/shim/opencv2/flann/flann.hpp" -nt "$o" ]; then
    $1 -c "$REF/$2" -o "$o"
  fi
  OBJS="$OBJS $o"
}
for f in detectors/helpers.cpp detectors/affinedetectors/pyramid.cpp detectors/affinedetectors/affine.cpp \
         detectors/affinedetectors/scale-space-detector.cpp matching/siftdesc.cpp synth-detection.cpp \
         detectors/mser/extrema/extrema.cpp detectors/mser/extrema/sortPixels.cpp detectors/mser/extrema/getExtrema.cpp \
         detectors/mser/extrema/libExtrema.cpp detectors/mser/extrema/boundary.cpp detectors/mser/extrema/suballoc.cpp \
         detectors/mser/extrema/optThresh.cpp detectors/mser/extrema/preprocess.cpp detectors/mser/utls/matrix.cpp \
         matching/matching.cpp; do
  cc_one "g++ $CXXFLAGS" $f
done
for f in detectors/mser/utls/timeutls.c detectors/mser/LL/LL.c detectors/mser/LL/LLstr.c detectors/mser/LL/LLio.c \
         detectors/mser/LL/LLfile.c detectors/mser/LL/LLmergeSort.c; do
  cc_one "gcc -O3 -fPIC -w -DA64 -I$REF/detectors/mser/LL -I$REF/detectors/mser/utls" $f
done
for f in degensac/DegUtils.c degensac/exp_ranF.c degensac/exp_ranH.c degensac/Ftools.c degensac/hash.c degensac/Htools.c \
         degensac/ranF.c degensac/ranH.c degensac/rtools.c degensac/utools.c degensac/lapwrap.c \
         matutls/minv.c matutls/trnm.c matutls/mmul.c matutls/mattr.c matutls/rmmult.c matutls/svduv.c matutls/qrbdv.c \
         matutls/ldumat.c matutls/ldvmat.c matutls/atou1.c matutls/atovm.c; do
  cc_one "gcc $CFLAGS" $f
done
g++ $CXXFLAGS -c $HERE/ref_api.cpp -o $OUT/obj/ref_api.o
g++ $CXXFLAGS -I$REF/degensac -c $HERE/ref_api_matching.cpp -o $OUT/obj/ref_api_matching.o
g++ -shared -fopenmp -o $OUT/libmods_ref.so $OBJS $OUT/obj/ref_api.o $OUT/obj/ref_api_matching.o $BLAS -Wl,-rpath,$BLASDIR -lm
echo "build_ref: built $OUT/libmods_ref.so"
