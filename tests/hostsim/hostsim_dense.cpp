// TEST INFRASTRUCTURE ONLY -- see hostsim_api.cpp.
#include "../../ilqc_b200/csrc/dense_kernels.inl"
