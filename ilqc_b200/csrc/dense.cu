// Generic dense small-system kernels (see dense_kernels.inl).
#include "dense_kernels.inl"
