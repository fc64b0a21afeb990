// Cluster GEMM kernels whose epilogue pushes the TP reduce-scatter words over NVLink (b200_wo_gemm_rs; see gemm_cluster_inst.cuh).
#define B200_GEMM_VAR 1
#define B200_GEMM_LAUNCH_NAME launch_cluster_gemm_rs
#include "gemm_cluster_inst.cuh"
