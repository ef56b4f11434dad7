/* libi8ie_sm100.so — C ABI of the B200 (sm_100a) backend for the i8ie INT8 hot path.
 *
 * This is the drop-in boundary: it replaces the reference's L0+L1 for the path
 * (Intel MKL CBLAS + the OpenMP loops around it, /root/reference include/layer.h:3)
 * underneath the unchanged `i8ie` Python surface. The only C ABI the reference
 * itself crosses on this path is
 *     cblas_gemm_s8u8s32(RowMajor, NoTrans, Trans, RowOffset, m, n, k, 1, A_u8, k, 0,
 *                        B_s8, k, 0, 0, C_s32, n, oc)          conv2d.cc:131-133
 *                                                              fully_connected.cc:39-41
 * with caller-owned buffers and no error return; every entry point below cites
 * the reference code it replaces (file:line in /root/reference).
 *
 * Conventions
 *  - plain C: pointers, sizes, scalars. No C++/torch types, no exceptions.
 *  - every pointer is a DEVICE pointer unless the name ends in _host.
 *  - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it.
 *  - return 0 on success, a negative I8IE_E* code otherwise; i8ie_last_error()
 *    returns a thread-local message for the last failure.
 *  - the caller owns every buffer; plan objects own only their own descriptors /
 *    packed-weight copies and are released with the matching *_destroy.
 *  - activations (u8) are NHWC with a channel pitch `cp` that is a multiple of 16
 *    bytes (pad lanes carry the tensor's zero_point); 2-D activations are
 *    row-major with a row pitch that is a multiple of 16 bytes. fp32 tensors are
 *    dense NCHW as in the reference (conv2d.cc:64-67).
 *  - arithmetic contract (bit-exact vs the reference): SURVEY.md Appendix A.
 */
#ifndef I8IE_SM100_H
#define I8IE_SM100_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define I8IE_API __attribute__((visibility("default")))
#else
#define I8IE_API
#endif

#define I8IE_OK 0
#define I8IE_EINVAL (-1)   /* bad argument / unsupported shape */
#define I8IE_ECUDA (-2)    /* CUDA runtime / driver error */
#define I8IE_ENOTSM100 (-3) /* device is not compute capability 10.0 */

#define I8IE_EPI_RELU 1     /* fuse relu<u8> (functional.cc:15-26): y = max(y, zp_out) */

/* Version / diagnostics */
I8IE_API const char* i8ie_last_error(void);
I8IE_API const char* i8ie_version(void);
I8IE_API int i8ie_device_check(void);          /* 0 iff the current device is sm_100 */
/* Number of kernels this library has launched in this process (gpu_launches in bench.py). */
I8IE_API int64_t i8ie_launch_count(void);

/* ---- element-wise kernels (HBM-bound) ------------------------------------ */

/* A1: quantize(Tensor<float>&, scale, zp), quantize_utils.cc:44-52 (unclamped, trunc):
 *   q[i] = (u8)(int)(x[i] / scale + zp). Flat, dense. */
I8IE_API int i8ie_quantize_f32_u8(const float* x, uint8_t* q, int64_t n, float scale, int zp, void* stream);

/* Same arithmetic, fused with the NCHW(f32) -> NHWC(u8, pitch cp) re-layout the conv
 * kernels consume; pad lanes [c, cp) are written with zp. */
I8IE_API int i8ie_quantize_nchw_f32_nhwc_u8(const float* x, uint8_t* q, int n, int c, int h, int w, int cp,
                                   float scale, int zp, void* stream);

/* "_indirect" variants: identical arithmetic, but the SOURCE ADDRESS is read from device memory
 * (*x_slot) when the kernel runs. A forward captured once into a CUDA graph can then be replayed
 * on any input buffer by writing its address into the slot (stream-ordered) — no staging copy.
 * The address stored in the slot must be 16-byte aligned. (No reference counterpart: the
 * reference passes the Tensor by reference, module.py:20.) */
I8IE_API int i8ie_quantize_f32_u8_indirect(const float* const* x_slot, uint8_t* q, int64_t n, float scale, int zp,
                                  void* stream);
I8IE_API int i8ie_quantize_nchw_f32_nhwc_u8_indirect(const float* const* x_slot, uint8_t* q, int n, int c, int h,
                                            int w, int cp, float scale, int zp, void* stream);
/* dst[0, nbytes) = (*src_slot)[0, nbytes); nbytes % 16 == 0, both 16-byte aligned. */
I8IE_API int i8ie_copy_indirect(const void* const* src_slot, void* dst, int64_t nbytes, void* stream);

/* ---- multi-GPU result exchange (no reference counterpart: the reference is single-process) ----
 * Batch shards are independent; the only exchange is the result. Each rank packs
 *   chunk = { int64 agree; int32 rows; int32 cols; float logits[rows_cap * cols] }
 * (agree = rows whose first-maximum argmax equals ref_argmax[r]; ref_argmax may be NULL), ONE
 * NCCL all-gather moves the chunks, and unpack writes the dense [sum rows, cols] logits in rank
 * order plus the summed count. chunk bytes = i8ie_top1_chunk_bytes(rows_cap, cols), a multiple
 * of 16 when rows_cap * cols is a multiple of 4 (pad rows_cap otherwise). */
I8IE_API int64_t i8ie_top1_chunk_bytes(int rows_cap, int cols);
I8IE_API int i8ie_top1_pack(const float* logits, const int64_t* ref_argmax, int rows, int cols, void* packed,
                   void* stream);
I8IE_API int i8ie_top1_unpack(const void* gathered, int world, int64_t chunk_bytes, float* logits_all,
                     int64_t* agree_total, void* stream);

/* ---- the same exchange over NVLink / NVSwitch PEER MEMORY (csrc/peer_exchange.cu) --------------
 * No NCCL call and no host involvement per step, so forward + exchange replay as ONE CUDA graph.
 * Every rank owns one shareable buffer of i8ie_peer_exchange_bytes(world, chunk) bytes
 *   [ flags: 16 x u64 | gathered[2][world][chunk] ]        (zero-initialised by i8ie_peer_alloc)
 * created with i8ie_peer_alloc (cudaMalloc + cudaIpcGetMemHandle; handle64 = the 64-byte IPC handle
 * to hand to the peers, e.g. through torch.distributed.all_gather_object) and mapped by every peer
 * with i8ie_peer_open. peer_bases[p] = rank p's buffer as mapped in the calling process (own
 * buffer: the pointer i8ie_peer_alloc returned).
 *   i8ie_top1_pack_push   packs this rank's chunk (as i8ie_top1_pack) straight into slot `rank` of
 *                         gathered[seq & 1] on EVERY rank with peer stores, then publishes seq in word
 *                         `rank` of every rank's flags (st.release.sys). seq = ++*seq_counter, a u64
 *                         in local device memory (initially 0), so a graph replay needs no argument.
 *   i8ie_top1_wait_unpack waits (bounded, ~15 s) until all `world` flag words of the local buffer
 *                         reached *seq_counter, then unpacks gathered[seq & 1] as i8ie_top1_unpack does.
 *                         A timeout is reported through i8ie_tc_error_poll (code 7).
 * One rank per GPU; all ranks must call the pair once per step, in the same order. */
I8IE_API int64_t i8ie_peer_exchange_bytes(int world, int64_t chunk_bytes);
I8IE_API int i8ie_peer_alloc(int64_t bytes, void** ptr, void* handle64);
I8IE_API int i8ie_peer_open(const void* handle64, void** ptr);
I8IE_API int i8ie_peer_close(void* ptr);
I8IE_API int i8ie_peer_free(void* ptr);
I8IE_API int i8ie_top1_pack_push(const float* logits, const int64_t* ref_argmax, int rows, int cols,
                                 void* const* peer_bases, int world, int rank, int64_t chunk_bytes,
                                 void* seq_counter, void* stream);
I8IE_API int i8ie_top1_wait_unpack(const void* mine, int world, int64_t chunk_bytes, const void* seq_counter,
                                   float* logits_all, int64_t* agree_total, void* stream);

/* ---- F1: the FP32 forward (calibration side; feeds Calibrator::sample) -------------------------
 * Conv2d::forward_prop(Tensor<float>&&), conv2d.cc:63-98: im2col + cblas_sgemm + bias, here one
 * implicit-GEMM fp32 FMA kernel. x dense NCHW [n,c,h,w], w OIHW [kc,c,kh,kw], bias [kc],
 * y dense NCHW [n,kc,oh,ow], oh = (h - kh + 2*pad)/stride + 1 (conv2d.cc:71-72).
 * minmax2 (device float[2], may be NULL): {min, max} of everything written is folded into it with
 * atomics — the caller initialises it to {+FLT_MAX, -FLT_MAX}; this is the calibrator's range
 * (conv2d.cc:94-96 -> calibrator.cc:6-27 at quantile 1) without a second pass over y.
 * Parity is tolerance-only (fp32 summation order differs from MKL's, as any two sgemms do). */
I8IE_API int i8ie_conv2d_f32(const float* x, const float* w, const float* bias, float* y, int n, int c, int h,
                             int wd, int kc, int kh, int kw, int stride, int pad, float* minmax2, void* stream);
/* Linear::forward_prop(Tensor<float>&&), fully_connected.cc:5-21: y[m,n] = x[m,k] . w[n,k]^T + bias. */
I8IE_API int i8ie_linear_f32(const float* x, const float* w, const float* bias, float* y, int m, int n, int k,
                             float* minmax2, void* stream);
/* relu<float>, functional.cc:5-13: y = x > 0 ? x : 0. */
I8IE_API int i8ie_relu_f32(const float* x, float* y, int64_t n, void* stream);
/* max_pool2d<float>, functional.cc:36-64: dense NCHW, no padding, floor output size. */
I8IE_API int i8ie_maxpool_f32_nchw(const float* x, float* y, int n, int c, int h, int w, int ksize, int stride,
                                   void* stream);

/* A5: dequantize(float*, u8*, size, scale, zp), quantize_utils.cc:38-42:
 *   x[i] = (float)((int)q[i] - zp) * scale. Flat, dense. */
I8IE_API int i8ie_dequantize_u8_f32(const uint8_t* q, float* x, int64_t n, float scale, int zp, void* stream);
/* Strided rows variant: q is [rows, pitch] (pitch >= cols), x dense [rows, cols]. */
I8IE_API int i8ie_dequantize_rows_u8_f32(const uint8_t* q, float* x, int rows, int cols, int pitch,
                                float scale, int zp, void* stream);

/* A4: down_scale, quantize_utils.cc:27-36 (standalone requantise s32 -> u8):
 *   d = ((float)acc*sa)*sb; r = d/sc + zp_c; y = r>=255 ? 255 : r<0 ? 0 : (u8)trunc(r). */
I8IE_API int i8ie_downscale_s32_u8(const int32_t* acc, uint8_t* y, int64_t n, float sa, float sb, float sc,
                          int zp_c, void* stream);

/* A9 (reduction half): min/max of an fp32 buffer — what Calibrator::sample
 * (calibrator.cc:6-23) + the sort in get_range (:25-27) yield for quantile 1 when
 * every value is kept. out2 = {min, max} (device, 2 floats); workspace must hold
 * i8ie_minmax_workspace_bytes() bytes and be zero-initialised once. */
I8IE_API int64_t i8ie_minmax_workspace_bytes(void);
I8IE_API int i8ie_minmax_f32(const float* x, int64_t n, float* out2, void* workspace, void* stream);
/* A9 (scalar half): calibrator.cc:28-35 with (min,max) as input. Host function. */
I8IE_API int i8ie_range_from_minmax_host(float mn, float mx, float* scale_host, uint8_t* zp_host);

/* A10: relu<u8_t>, functional.cc:15-26: y = max(x, zp). Flat. */
I8IE_API int i8ie_relu_u8(const uint8_t* x, uint8_t* y, int64_t n, int zp, void* stream);

/* A11: max_pool2d<u8_t>, functional.cc:36-64 (no padding, floor): NHWC in (pitch cp)
 * -> NHWC out (pitch cp), or, if out_nchw != 0, dense NCHW out (the flatten order
 * x.reshape(-1, c*oh*ow) needs, tensor.h:106-133). */
I8IE_API int i8ie_maxpool_u8_nhwc(const uint8_t* x, uint8_t* y, int n, int h, int w, int c, int cp,
                         int ksize, int stride, int out_nchw, void* stream);
/* The same pooling into a PHYSICALLY PADDED NHWC tensor y[n][oh + 2*out_pad][ow + 2*out_pad][out_cp]:
 * border pixels hold pad_value (the tensor's zero point: conv2d.cc:24-25 pads im2col with in.zero_point),
 * channel pitch out_cp (c <= out_cp <= cp, multiple of 16). Feeds the "row mode" convolution plans
 * (i8ie_conv2d_plan_create impl 4). ksize = stride = 1 is a plain pad-copy of an NHWC tensor. */
I8IE_API int i8ie_maxpool_u8_nhwc_padded(const uint8_t* x, uint8_t* y, int n, int h, int w, int c, int cp,
                                         int ksize, int stride, int out_cp, int out_pad, int pad_value,
                                         void* stream);

/* Layout glue for the NCHW-facing API (tensor.h:40-47 / pybind11.cc:14-15). */
I8IE_API int i8ie_u8_nchw_to_nhwc(const uint8_t* x, uint8_t* y, int n, int c, int h, int w, int cp, int pad_value, void* stream);
I8IE_API int i8ie_u8_nhwc_to_nchw(const uint8_t* x, uint8_t* y, int n, int c, int h, int w, int cp, void* stream);

/* ---- weight preparation (once per model) ---------------------------------- */

/* A8: quantize_weight, layer.cc:6-26. Host function on host buffers (runs once at
 * convert()): shared min/max over weight U bias, scale=(max-min)/127, truncating
 * unclamped casts. Returns the scale through *scale_host. */
I8IE_API int i8ie_quantize_weight_host(const float* w_host, int64_t nw, const float* b_host, int64_t nb,
                              int8_t* qw_host, int8_t* qb_host, float* scale_host);

/* A2b / A3: zero-point + bias offsets, conv2d.cc:117-124 / fully_connected.cc:30-38.
 * qw is the s8 weight in the REFERENCE's row order [n, K] (OIHW flattened), so the
 * sequential fp32 accumulation of t is reproduced exactly:
 *   t[j]  = sum_k (float)(in_zp * qw[j][k])            (fp32, k ascending)
 *   conv: oc[j] = (int)((float)qb[j] / in_scale - t[j]);  bias_f[j] = 0
 *   fc:   oc[j] = (int)(-t[j]);                          bias_f[j] = (float)qb[j] / in_scale
 * oc and bias_f have n entries (device). */
I8IE_API int i8ie_zp_offsets(const int8_t* qw, const int8_t* qb, int n, int k, int in_zp, float in_scale,
                    int is_conv, int32_t* oc, float* bias_f, void* stream);

/* Repack OIHW s8 weights into the kernels' K-major [kc_pad, kh, kw, cp] layout
 * (zero in every pad lane / pad row). */
I8IE_API int i8ie_pack_conv_weight(const int8_t* qw_oihw, int8_t* w_packed, int kc, int c, int kh, int kw,
                          int kc_pad, int cp, void* stream);

/* ---- the GEMM-shaped ops (tensor-core bound) -------------------------------- */

typedef struct i8ie_conv_plan i8ie_conv_plan;

/* Plan for Conv2d::forward_prop(Tensor<u8>&&), conv2d.cc:100-142, for one input
 * geometry. w_packed is [kc_pad, kh, kw, cp] s8 (i8ie_pack_conv_weight) and must
 * outlive the plan. x is NHWC u8 with pitch cp; y is NHWC u8 with pitch
 * out_cp (multiple of 16, >= kc). Replaces im2col (conv2d.cc:34-49), the
 * per-image cblas_gemm_s8u8s32 (:131-133), down_scale (:134-135) and transpose (:136).
 * impl: 0 = auto, 1 = force the SIMT dp4a kernel, 2 = force tcgen05 (error if ineligible),
 * 4 = tcgen05 ROW MODE: x is then the PHYSICALLY PADDED tensor [n][h + 2*pad][w + 2*pad][cp] whose border
 * holds in.zero_point (i8ie_maxpool_u8_nhwc_padded writes it) followed by >= 128 readable bytes, and cp must
 * be i8ie_conv2d_row_mode_cp(...) (!= 0). The GEMM K index is (filter row, byte of the contiguous kw * cp
 * run) instead of (tap, channel padded to 128): fewer K blocks when c is not a multiple of 128. */
I8IE_API i8ie_conv_plan* i8ie_conv2d_plan_create(int n, int c, int h, int w, int cp, int kc, int kh, int kw,
                                        int stride, int pad, int out_cp, const int8_t* w_packed,
                                        int kc_pad, int impl);
I8IE_API void i8ie_conv2d_plan_destroy(i8ie_conv_plan* plan);
/* Channel pitch the padded input of a row-mode plan must have, or 0 when the layer should use the plain
 * plans (stride != 1, narrow N, or no K saved). cp_plain = the pitch the plain NHWC tensor would have. */
I8IE_API int i8ie_conv2d_row_mode_cp(int c, int cp_plain, int kh, int kw, int stride, int pad, int out_cp);
/* Which kernel the plan resolved to: 1 = SIMT dp4a, 2 = tcgen05 (TMA im2col), 3 = tcgen05 stem
 * (small-C strided first layer: the plan owns a bordered superpixel copy of the input), 4 = tcgen05 row mode. */
I8IE_API int i8ie_conv2d_plan_impl(const i8ie_conv_plan* plan);
/* y = requant(conv(x) + oc) [+relu]; sa=in.scale, sb=weight scale, sc=layer scale_,
 * zp_in = in.zero_point (spatial padding value, conv2d.cc:129-130), zp_out = layer
 * zero_point_. acc_out (optional, may be NULL) receives the s32 accumulators incl.
 * oc as [n*oh*ow, kc] for parity tests. */
I8IE_API int i8ie_conv2d_u8(i8ie_conv_plan* plan, const uint8_t* x, uint8_t* y, const int32_t* oc,
                   float sa, float sb, float sc, int zp_in, int zp_out, int flags,
                   int32_t* acc_out, void* stream);

/* F4 extension (opt-in; NOT in the reference, whose weight scale is one per tensor, layer.cc:18-19):
 * per-output-channel weight scales. sb_vec = device float[kc] that must outlive the plan's use;
 * [sb_min, sb_max] bound its entries (the exact fast requantise path is guarded on both). Every later
 * i8ie_conv2d_u8 / i8ie_conv2d_f32_u8 call of the plan requantises channel n with
 *   d = ((float)acc * sa) * sb_vec[n]   (otherwise quantize_utils.cc:27-36 unchanged)
 * and ignores its scalar `sb`. sb_vec = NULL restores the per-tensor behaviour. */
I8IE_API int i8ie_conv2d_plan_set_channel_scales(i8ie_conv_plan* plan, const float* sb_vec, float sb_min,
                                                 float sb_max);

/* Module.__call__'s input quantise (i8ie/module.py:20 -> quantize_utils.cc:44-52) fused into
 * the first convolution: x is the fp32 NCHW image; q = (u8)(x/in_scale + in_zp) is produced
 * on the fly into the plan's stem buffer and never stored as an NHWC tensor. Only valid
 * for plans with impl 3; returns I8IE_EINVAL otherwise (quantise + i8ie_conv2d_u8 then). */
I8IE_API int i8ie_conv2d_f32_u8(i8ie_conv_plan* plan, const float* x_nchw, float in_scale, int in_zp, uint8_t* y,
                       const int32_t* oc, float sb, float sc, int zp_out, int flags, int32_t* acc_out,
                       void* stream);

/* Same with the image address read from *x_slot at run time (see the "_indirect" note above). */
I8IE_API int i8ie_conv2d_f32_u8_indirect(i8ie_conv_plan* plan, const float* const* x_slot, float in_scale,
                                int in_zp, uint8_t* y, const int32_t* oc, float sb, float sc, int zp_out,
                                int flags, int32_t* acc_out, void* stream);

/* Linear::forward_prop(Tensor<u8>&&), fully_connected.cc:22-52:
 *   acc = x[m,k] * w[n,k]^T + oc[n];  acc = (int)((float)acc + bias_f[n]);  y = requant(acc) [+relu]
 * x pitch = ldx, w is [n_pad, ldw] s8 K-major (pad rows/lanes zero), y pitch = ldy.
 * Dispatches on m only (small-m weight-streaming kernel vs tiled kernel). */
I8IE_API int i8ie_fc_u8(const uint8_t* x, int ldx, const int8_t* w, int ldw, int n_pad, uint8_t* y, int ldy,
               int m, int n, int k, const int32_t* oc, const float* bias_f, float sa, float sb,
               float sc, int zp_out, int flags, int32_t* acc_out, int impl, void* stream);
/* i8ie_fc_u8 with per-output-channel weight scales (F4 extension, see
 * i8ie_conv2d_plan_set_channel_scales): sb_vec = device float[n], bounded by [sb_min, sb_max]. */
I8IE_API int i8ie_fc_u8_pc(const uint8_t* x, int ldx, const int8_t* w, int ldw, int n_pad, uint8_t* y, int ldy,
                           int m, int n, int k, const int32_t* oc, const float* bias_f, float sa,
                           const float* sb_vec, float sb_min, float sb_max, float sc, int zp_out, int flags,
                           int32_t* acc_out, int impl, void* stream);

/* Small-batch fully_connected.cc:39-41 is a weight stream from HBM. A tiled box of 128-byte rows that lie
 * `ldw` bytes apart streams at ~2.3 TB/s on B200; the same bytes stored as contiguous 16 KB blocks stream at
 * ~6 TB/s (tools/ubench/i8_peak.cu). The owner of an fc weight buffer w [n_pad][ldw] (both multiples of 128)
 * may therefore attach a second buffer of i8ie_fc_weight_tiled_bytes() bytes (16-byte aligned): attach
 * fills it with [n_pad/128][ldw/128] blocks of 128 rows x 128 bytes in the kernel's swizzled shared-memory
 * image and registers the pair; later i8ie_fc_u8* calls with this w read the weights through it when the
 * tcgen05 kernel runs with a 128- or 256-wide N tile (results identical). detach(w) must be called before
 * either buffer is freed or rewritten. i8ie_fc_weight_tiled_bytes returns 0 for shapes without a tiled form. */
I8IE_API int64_t i8ie_fc_weight_tiled_bytes(int n_pad, int ldw);
I8IE_API int i8ie_fc_weight_tiled_attach(const int8_t* w, int n_pad, int ldw, int8_t* w_tiled, void* stream);
I8IE_API int i8ie_fc_weight_tiled_detach(const int8_t* w);

/* Linear::forward_prop(Tensor<u8>&&) (fully_connected.cc:22-52) of a model's LAST layer together with
 * Module.__call__'s final dequantize (i8ie/module.py:22-24 -> quantize_utils.cc:54-58 -> :38-42):
 * y as i8ie_fc_u8 / i8ie_fc_u8_pc (sb_vec may be NULL: per-tensor scale sb), and
 *   deq_out[m, n] (dense fp32) = ((float)y - zp_out) * sc.
 * Classifier heads (n_pad == ldy == 16) do both in one kernel; other shapes run the fc kernel and
 * then the standalone dequantise — same values either way. */
I8IE_API int i8ie_fc_u8_deq(const uint8_t* x, int ldx, const int8_t* w, int ldw, int n_pad, uint8_t* y, int ldy,
                            int m, int n, int k, const int32_t* oc, const float* bias_f, float sa, float sb,
                            const float* sb_vec, float sb_min, float sb_max, float sc, int zp_out, int flags,
                            int impl, float* deq_out, void* stream);

/* Debug hook (not part of the reference-facing surface): synchronises the device and
 * returns the first protocol error (mbarrier wait timeout) a tensor-core kernel recorded
 * (0 = none, >0 = role that timed out), optionally clearing it; negative on CUDA errors. */
I8IE_API int i8ie_debug_tc_error(int reset);

/* Same flag, read from a host-mapped mirror WITHOUT any CUDA call or synchronisation: valid once
 * the work in question has been synchronised by the caller (e.g. right after a device-to-host
 * copy of a result). The Python layer checks it on every result read-back and raises, so a
 * pipeline fault can never return garbage activations with rc = 0. */
I8IE_API int i8ie_tc_error_poll(int reset);

#ifdef __cplusplus
}
#endif
#endif /* I8IE_SM100_H */
