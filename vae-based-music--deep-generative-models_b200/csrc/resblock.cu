// resblock.cu — vqb_resblock_fwd / vqb_resblock_bwd_data: the pre-activation residual block of
// ResnetConv1DBlock.call (resnet.py:11-18,29).  VQB_PREC_FP32 runs the exact CUDA-core contraction
// kernels of conv_fp32.cu (two launches, ReLU / bias / residual fused into load and epilogue);
// VQB_PREC_BF16 / TF32 run the fused tcgen05 kernel of resblock_tc.cu for the shapes it supports.
#include "common.cuh"

namespace vqb {
int conv1d_fwd_fp32(const vqb_conv_desc* d, const float* x, const float* w, const float* bias,
                    const float* residual, float* y, cudaStream_t st);
int conv1d_dgrad_fp32(const vqb_conv_desc* d, const float* dy, const float* w, const float* x,
                      const float* dx_add, float* dx, cudaStream_t st);
int resblock_fwd_tc(const vqb_resblock_desc* d, const float* x, const float* w1, const float* b1,
                    const float* w2, const float* b2, float* h, float* y, uint32_t* xbits, uint32_t* hbits, cudaStream_t st);
int resblock_bwd_tc(const vqb_resblock_desc* d, const float* x, const float* h, const float* dy, const float* w1,
                    const float* w2, float* dh, float* dx, const uint32_t* xbits, const uint32_t* hbits, cudaStream_t st);
bool resblock_tc_supported(const vqb_resblock_desc* d);
}  // namespace vqb

using namespace vqb;

static int check_rb(const vqb_resblock_desc* d) {
  VQB_REQUIRE(d != nullptr, "resblock desc is NULL");
  VQB_REQUIRE(d->B >= 0 && d->L >= 0 && d->C > 0 && d->F > 0 && d->dilation >= 1,
              "resblock desc: bad shape B=%d L=%d C=%d F=%d dil=%d", d->B, d->L, d->C, d->F, d->dilation);
  return VQB_OK;
}

extern "C" {

int vqb_resblock_supports(const vqb_resblock_desc* d) {
  if (!d) return 0;
  if (d->precision == VQB_PREC_FP32) return 1;
  return resblock_tc_supported(d) ? 1 : 0;
}

int vqb_resblock_fwd(const vqb_resblock_desc* d, const float* x, const float* w1, const float* b1,
                     const float* w2, const float* b2, float* h, float* y, void* stream) {
  VQB_ARCH();
  int rc = check_rb(d);
  if (rc) return rc;
  VQB_REQUIRE(x && w1 && w2 && y, "vqb_resblock_fwd: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  // h may be NULL on the tensor-core paths (inference: the intermediate is not needed outside the kernel and is not stored)
  if (d->precision != VQB_PREC_FP32) return resblock_fwd_tc(d, x, w1, b1, w2, b2, h, y, nullptr, nullptr, st);
  VQB_REQUIRE(h, "vqb_resblock_fwd: the fp32 path needs h (it is the input of its second kernel)");
  vqb_conv_desc c1{d->B, d->L, d->C, d->F, 3, 1, d->dilation, 1, VQB_PREC_FP32};
  rc = conv1d_fwd_fp32(&c1, x, w1, b1, nullptr, h, st);
  if (rc) return rc;
  vqb_conv_desc c2{d->B, d->L, d->F, d->C, 3, 1, 1, 1, VQB_PREC_FP32};
  return conv1d_fwd_fp32(&c2, h, w2, b2, x, y, st);
}

int vqb_resblock_bwd_data(const vqb_resblock_desc* d, const float* x, const float* h, const float* dy,
                          const float* w1, const float* w2, float* dh, float* dx, void* stream) {
  VQB_ARCH();
  int rc = check_rb(d);
  if (rc) return rc;
  VQB_REQUIRE(x && h && dy && w1 && w2 && dh && dx, "vqb_resblock_bwd_data: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->precision != VQB_PREC_FP32) return resblock_bwd_tc(d, x, h, dy, w1, w2, dh, dx, nullptr, nullptr, st);
  vqb_conv_desc c2{d->B, d->L, d->F, d->C, 3, 1, 1, 1, VQB_PREC_FP32};
  rc = conv1d_dgrad_fp32(&c2, dy, w2, h, nullptr, dh, st);  // dh = (h>0) * conv2^T(dy)
  if (rc) return rc;
  vqb_conv_desc c1{d->B, d->L, d->C, d->F, 3, 1, d->dilation, 1, VQB_PREC_FP32};
  return conv1d_dgrad_fp32(&c1, dh, w1, x, dy, dx, st);     // dx = (x>0) * conv1^T(dh) + dy
}

/* Sign-mask variants for the tensor-core precisions (resblock_tc.cu): the forward also writes xbits / hbits [B, L] (bit c of
   word t = channel c of x / h at t is > 0), the data gradient reads those 8 bytes per position instead of the two fp32
   tensors (x and h stay the operands of vqb_resblock_wgrad). */
int vqb_resblock_fwd_masks(const vqb_resblock_desc* d, const float* x, const float* w1, const float* b1, const float* w2,
                           const float* b2, float* h, float* y, uint32_t* xbits, uint32_t* hbits, void* stream) {
  VQB_ARCH();
  int rc = check_rb(d);
  if (rc) return rc;
  VQB_REQUIRE(x && w1 && w2 && h && y && xbits && hbits, "vqb_resblock_fwd_masks: NULL pointer");
  VQB_REQUIRE(d->precision != VQB_PREC_FP32 && resblock_tc_supported(d),
              "vqb_resblock_fwd_masks: tensor-core precisions and shapes only (vqb_resblock_supports)");
  return resblock_fwd_tc(d, x, w1, b1, w2, b2, h, y, xbits, hbits, (cudaStream_t)stream);
}

int vqb_resblock_bwd_data_masks(const vqb_resblock_desc* d, const uint32_t* xbits, const uint32_t* hbits, const float* dy,
                                const float* w1, const float* w2, float* dh, float* dx, void* stream) {
  VQB_ARCH();
  int rc = check_rb(d);
  if (rc) return rc;
  VQB_REQUIRE(xbits && hbits && dy && w1 && w2 && dh && dx, "vqb_resblock_bwd_data_masks: NULL pointer");
  VQB_REQUIRE(d->precision != VQB_PREC_FP32 && resblock_tc_supported(d),
              "vqb_resblock_bwd_data_masks: tensor-core precisions and shapes only (vqb_resblock_supports)");
  return resblock_bwd_tc(d, nullptr, nullptr, dy, w1, w2, dh, dx, xbits, hbits, (cudaStream_t)stream);
}

/* both weight gradients of the block: dw1 / db1 from (ReLU(x), dh, dilation), dw2 / db2 from (ReLU(h), dy, 1).  The tensor-core
   precisions run them as ONE launch (wgrad_tc.cu: two problems share the grid), fp32 as two vqb_conv1d_wgrad calls. */
size_t vqb_resblock_wgrad_workspace_bytes(const vqb_resblock_desc* d) {
  if (!d) return 0;
  vqb_conv_desc c1{d->B, d->L, d->C, d->F, 3, 1, d->dilation, 1, d->precision};
  vqb_conv_desc c2{d->B, d->L, d->F, d->C, 3, 1, 1, 1, d->precision};
  if (d->precision != VQB_PREC_FP32 && d->C == d->F && wgrad_tc_supported(&c1)) return resblock_wgrad_tc_workspace_bytes(&c1, 1);
  c1.precision = c2.precision = VQB_PREC_FP32;  // no tensor-core weight-gradient kernel for this shape / mode (e.g. tf32)
  const size_t a = (vqb_conv1d_wgrad_workspace_bytes(&c1) + 255) & ~(size_t)255;
  return a + vqb_conv1d_wgrad_workspace_bytes(&c2);
}

int vqb_resblock_wgrad(const vqb_resblock_desc* d, const float* x, const float* h, const float* dy, const float* dh,
                       float* dw1, float* db1, float* dw2, float* db2, void* workspace, size_t workspace_bytes,
                       void* stream) {
  VQB_ARCH();
  int rc = check_rb(d);
  if (rc) return rc;
  VQB_REQUIRE(x && h && dy && dh && dw1 && dw2, "vqb_resblock_wgrad: NULL pointer");
  vqb_conv_desc c1{d->B, d->L, d->C, d->F, 3, 1, d->dilation, 1, d->precision};
  vqb_conv_desc c2{d->B, d->L, d->F, d->C, 3, 1, 1, 1, d->precision};
  if (d->precision != VQB_PREC_FP32 && d->C == d->F && wgrad_tc_supported(&c1) && d->B > 0 && d->L > 0)
    return resblock_wgrad_tc(&c1, 1, &d->dilation, &x, &h, &dy, &dh, &dw1, &db1, &dw2, &db2, workspace, workspace_bytes, (cudaStream_t)stream);
  c1.precision = c2.precision = VQB_PREC_FP32;  // shapes / modes without a tensor-core weight-gradient kernel: exact fp32
  const size_t a = (vqb_conv1d_wgrad_workspace_bytes(&c1) + 255) & ~(size_t)255;
  const size_t need = a + vqb_conv1d_wgrad_workspace_bytes(&c2);
  if (!workspace || workspace_bytes < need)
    return set_err(VQB_ERR_WORKSPACE, "vqb_resblock_wgrad workspace: need %zu bytes, got %zu", need, workspace_bytes);
  rc = vqb_conv1d_wgrad(&c1, x, dh, dw1, db1, workspace, a, stream);
  if (rc) return rc;
  return vqb_conv1d_wgrad(&c2, h, dy, dw2, db2, (char*)workspace + a, workspace_bytes - a, stream);
}

/* the same for n blocks of one shape (1 <= n <= 4; e.g. the four blocks of a DilatedResnet1D once its data-gradient chain is
   through): ONE launch on the tensor-core paths.  d->dilation is ignored, dilations[i] is block i's. */
size_t vqb_resblock_wgrad_batch_workspace_bytes(const vqb_resblock_desc* d, int32_t n) {
  if (!d || n < 1) return 0;
  vqb_conv_desc c1{d->B, d->L, d->C, d->F, 3, 1, 1, 1, d->precision};
  if (d->precision != VQB_PREC_FP32 && d->C == d->F && wgrad_tc_supported(&c1) && n <= 4) return resblock_wgrad_tc_workspace_bytes(&c1, n);
  return (size_t)n * ((vqb_resblock_wgrad_workspace_bytes(d) + 255) & ~(size_t)255);
}

int vqb_resblock_wgrad_batch(const vqb_resblock_desc* d, int32_t n, const int32_t* dilations, const float* const* x,
                             const float* const* h, const float* const* dy, const float* const* dh, float* const* dw1,
                             float* const* db1, float* const* dw2, float* const* db2, void* workspace,
                             size_t workspace_bytes, void* stream) {
  VQB_ARCH();
  int rc = check_rb(d);
  if (rc) return rc;
  VQB_REQUIRE(n >= 1 && dilations && x && h && dy && dh && dw1 && db1 && dw2 && db2, "vqb_resblock_wgrad_batch: bad arguments");
  vqb_conv_desc c1{d->B, d->L, d->C, d->F, 3, 1, 1, 1, d->precision};
  if (d->precision != VQB_PREC_FP32 && d->C == d->F && wgrad_tc_supported(&c1) && d->B > 0 && d->L > 0 && n <= 4)
    return resblock_wgrad_tc(&c1, n, dilations, x, h, dy, dh, dw1, db1, dw2, db2, workspace, workspace_bytes, (cudaStream_t)stream);
  const size_t per = (vqb_resblock_wgrad_workspace_bytes(d) + 255) & ~(size_t)255;
  if (!workspace || workspace_bytes < per * (size_t)n)
    return set_err(VQB_ERR_WORKSPACE, "vqb_resblock_wgrad_batch workspace: need %zu bytes, got %zu", per * (size_t)n, workspace_bytes);
  for (int i = 0; i < n; ++i) {
    vqb_resblock_desc di = *d;
    di.dilation = dilations[i];
    rc = vqb_resblock_wgrad(&di, x[i], h[i], dy[i], dh[i], dw1[i], db1[i], dw2[i], db2[i], (char*)workspace + per * i, per, stream);
    if (rc) return rc;
  }
  return VQB_OK;
}

}  // extern "C"
