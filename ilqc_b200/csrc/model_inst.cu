// One translation unit per model: nvcc -DILQC_MODEL_INST=<k> (see __graft_entry__.build / Makefile).
#include "model_kernels.inl"
