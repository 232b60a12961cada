// oracle/fma_probe.cu -- TEST INFRASTRUCTURE ONLY.
//
// The three squared-distance expressions of the tokenizer path, written in the SOURCE FORM of the CUDA packages the
// reference calls, so that `nvcc -ptx` (default -fmad=true, as those packages are built) shows which multiplies the
// compiler contracts into FMAs and in which order.  oracle/cpu_ref.c writes exactly that order out with fmaf(), and
// the act_b200 kernels with __fmaf_rn / __fmul_rn; tests/test_fma_probe.py compiles this file, rebuilds each result's
// expression tree from the PTX and compares it with the order the oracle states.
//
//   probe_fps      pointnet2_ops sampling_gpu.cu (not vendored; called from /root/reference/utils/misc.py:44):
//                      mag = x2*x2 + y2*y2 + z2*z2;   d = (x2-x1)*(x2-x1) + (y2-y1)*(y2-y1) + (z2-z1)*(z2-z1)
//   probe_knn      KNN_CUDA 0.2 knn.cu (not vendored; /root/reference/models/dvae.py:23,159): accumulate loop over the
//                      dimension: tmp = A[k] - B[k]; ssd += tmp*tmp
//   probe_chamfer  /root/reference/extensions/chamfer_dist/chamfer.cu:42-46 (in tree): x2 = b.x - a.x ...;
//                      d = x2*x2 + y2*y2 + z2*z2
#include <cuda_runtime.h>

extern "C" __global__ void probe_fps(const float *p1, const float *p2, float *out) {
    const float x1 = p1[0], y1 = p1[1], z1 = p1[2];
    const float x2 = p2[0], y2 = p2[1], z2 = p2[2];
    const float mag = x2 * x2 + y2 * y2 + z2 * z2;
    const float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
    out[0] = mag;
    out[1] = d;
}

extern "C" __global__ void probe_knn(const float *A, const float *B, float *out) {
    float ssd = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float tmp = A[k] - B[k];
        ssd += tmp * tmp;
    }
    out[0] = ssd;
}

extern "C" __global__ void probe_chamfer(const float *a, const float *b, float *out) {
    const float x1 = a[0], y1 = a[1], z1 = a[2];
    const float x2 = b[0] - x1, y2 = b[1] - y1, z2 = b[2] - z1;
    const float d = x2 * x2 + y2 * y2 + z2 * z2;
    out[0] = d;
}
