// TEST INFRASTRUCTURE ONLY -- see hostsim_api.cpp.  Compiled once per model with -DILQC_MODEL_INST=<k>.
#include "../../ilqc_b200/csrc/model_kernels.inl"
