// Single translation unit of libmods_b200.so (keeps the __constant__ tables in one definition and
// lets nvcc inline across kernel files).  Build: see mods_b200/build.py / Makefile.
#include "pyramid.cu"
#include "affine.cu"
#include "orient.cu"
#include "describe.cu"
#include "nn.cu"
#include "ransac.cu"
#include "mser.cu"
#include "synth.cu"
#include "capi.cu"
#include "ransac_host.cu"
#include "ransac_f_host.cu"
