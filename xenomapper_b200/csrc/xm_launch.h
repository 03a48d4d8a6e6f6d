/* xm_launch.h -- host-callable launchers of the kernels in xm_kernels.cu */
#pragma once
#include <cuda_runtime.h>
#include "xm_common.h"

namespace xm {
uint32_t tile_bytes(bool small);
cudaError_t launch_scan(const ScanArgs &a, bool small, cudaStream_t st);
cudaError_t launch_scan2(const ScanArgs &a, cudaStream_t st);      /* warp-autonomous scan, clean inputs only (xm_scan2.cuh) */
uint32_t scan2_tile_bytes();
uint32_t span_tile_bytes_min();
cudaError_t launch_classify2(const ClassifyArgs &a, cudaStream_t st);  /* the same for the primary stream (no errors, nothing to normalise) */
cudaError_t launch_classify(const ClassifyArgs &a, bool small, cudaStream_t st);
cudaError_t launch_size(const EmitArgs &a, cudaStream_t st);        /* the walk over rows (xm_emit.cuh) */
cudaError_t launch_prefix(const EmitArgs &a, cudaStream_t st);
cudaError_t launch_emit(const EmitArgs &a, cudaStream_t st);
cudaError_t launch_add_u64(unsigned long long *p, unsigned long long d, size_t n, cudaStream_t st);
cudaError_t launch_fill_u64(unsigned long long *p, unsigned long long v, size_t n, cudaStream_t st);
}  // namespace xm
