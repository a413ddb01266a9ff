/*
 * libb200mm -- C ABI of the B200-native (sm_100a) kernels behind the task-2C multimodal classifier hot path.
 *
 * The reference (KevinMathewT/multimodal-propaganda-meme-classification) is pure Python: it has no FFI of its own.
 * Its hot path bottoms out in ATen / cuBLASLt / cuDNN calls made by transformers and torchvision; each entry point
 * below names the reference call site(s) it replaces.  Paths are relative to the reference repository;
 * `$TF` = transformers/models/distilbert/modeling_distilbert.py, `$TV` = torchvision/models/resnet.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (CUDA), plain sizes, no torch types, no ownership transfer;
 *   - `stream` is a cudaStream_t; all work is enqueued asynchronously on it (CUDA-graph capturable);
 *   - return value: 0 = ok, > 0 = cudaError_t, < 0 = B200MM_ERR_* below; nothing throws across the boundary;
 *   - bf16 matrices are row-major with the innermost dimension contiguous; `ld*` are row strides in ELEMENTS;
 *   - alignment contract: bf16 pointers 16-byte aligned, row strides multiples of 8 elements.
 * There is no CPU implementation behind any of these symbols.
 */
#ifndef B200MM_H_
#define B200MM_H_

#ifdef __cplusplus
extern "C" {
#endif

#define B200MM_OK 0
#define B200MM_ERR_BAD_ARG (-1)   /* shape / alignment contract violated */
#define B200MM_ERR_NO_DRIVER (-2) /* cuTensorMapEncodeTiled could not be resolved */
#define B200MM_ERR_TENSORMAP (-3) /* the driver rejected a TMA tensor map */
#define B200MM_ERR_NOT_SM100 (-4) /* current device is not compute capability 10.x */
#define B200MM_JPEG_UNSUPPORTED (-10) /* a valid JPEG whose coding the split decoder does not handle */
#define B200MM_JPEG_CORRUPT (-11)     /* not a JPEG, damaged or truncated */

int b200mm_version(void);
int b200mm_num_sms(void);

/* ---- tcgen05 / TMA bf16 GEMM with fused epilogues ---------------------------------------------------------------
 * D[M,N] = A x B, fp32 accumulation in TMEM.  a_mn = 0: A stored [M,K]; 1: A stored [K,M].  b_mn = 0: B stored
 * [N,K]; 1: B stored [K,N].  epi: 0 store bf16 (acc + bias, dropout(p_drop, seed), + residual) | 1 GELU: out = z,
 * out2 = gelu(z) | 2 dGELU: out = acc * gelu'(aux) | 3 fp32 store | 4 fp32 atomic accumulate (split-K) | 5 ReLU.
 * Replaces: nn.Linear forward/backward behind q_lin/k_lin/v_lin/out_lin ($TF:187-189, :206), lin1/GELU/lin2
 * ($TF:223-227), the fusion head (example_scripts/Multimodal_example_task2C.txt:179, :184, :193), torchvision's
 * fc ($TV:206, :278) and every convolution of ResNet-50 ($TV:197, :118-130 via conv1x1 / conv3x3) as implicit GEMM. */
int b200mm_gemm_bf16(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, int M, int N,
                     int K, int epi, const float* bias, const void* residual, long long ldr, const void* aux,
                     long long ld_aux, void* out, long long ldc, void* out2, long long ld2, int splits, int block_n,
                     float p_drop, unsigned long long seed, float* col_stats, void* stream);
/* residual == out (same pointer and stride, epi 0, no dropout): out += A x B (+ bias) in place -- the epilogue issues
 * TMA reduce-add stores, the add happens in L2 (bf16), the SM never loads the residual.
 * col_stats (nullable; epi 0 without bias / residual / dropout): fp32 [2N], col_stats[n] += sum_m out[m,n],
 * col_stats[N+n] += sum_m out[m,n]^2 over the STORED bf16 values -- the train-mode BatchNorm statistics of a
 * convolution output come out of the convolution's own epilogue (feeds b200mm_batchnorm_fwd_stats). */

/* out = bf16(A x B + (mask bit ? residual : 0)): data gradient of a residual block's first convolution + the identity
 * branch's gradient (dout o ReLU mask) in one epilogue; mask = the 1-bit-per-element [M, N/8] ReLU mask written by
 * b200mm_batchnorm_fwd_stats.  Replaces autograd's add of the two branch gradients behind
 * torchvision/models/resnet.py:155-161 (out += identity; relu).  N % 32 == 0. */
int b200mm_gemm_bf16_maskres(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, int M, int N,
                             int K, const void* residual, long long ldr, const unsigned char* mask, long long ld_mask,
                             void* out, long long ldc, void* stream);

/* Dispatch knobs (the defaults are the measured optimum; for A/B micro-benchmarks and for tests that must reach both
 * sides of a dispatch decision): 0 = smallest reduction depth in 64-wide k blocks the CTA-pair GEMM takes (8),
 * 1 = B-resident GEMM mode for short unsplit K (1), 2 = largest tensor in MB whose BatchNorm backward runs as one
 * cooperative launch (0 = never, the default: it measured 0.7 % slower on the config-2 step). */
int b200mm_tune(int knob, int value);

/* Implicit-GEMM convolution on the same kernel: the activation operand is gathered from NHWC memory by im2col-mode
 * TMA loads (no im2col matrix exists).  x: bf16 [N,H,W,C], C % 64 == 0; w: bf16 OHWI-flattened [Cout, k*k*C].
 * conv_fwd also serves the stride-1 data gradient (call it on dY with the weight from conv_weight_rotate).
 * Replaces conv3x3 of torchvision/models/resnet.py:19-31, :118-130 (cuDNN fprop / dgrad / wgrad). */
int b200mm_conv_fwd(const void* x, int N, int H, int W, int C, const void* w, int Cout, int ksize, int stride, int pad,
                    int epi, const float* bias, const void* residual, long long ldr, void* out, long long ldc,
                    float* col_stats, void* stream);
int b200mm_conv_wgrad(const void* dy, long long ld_dy, const void* x, int N, int H, int W, int C, int Cout, int ksize,
                      int stride, int pad, float* dw, int splits, void* stream);
int b200mm_conv_weight_rotate(const void* w, void* w_rot, int Cout, int Cin, int ksize, void* stream);
/* every rotation of a backward pass in one launch; table: DEVICE int64 [n][4] = {w, w_rot, Cout << 32 | Cin, ksize^2} */
int b200mm_conv_weight_rotate_multi(const long long* table, int n, void* stream);

/* ---- fused attention (head_dim 64, S <= 512) -------------------------------------------------------------------
 * out[B*S, H*64] = softmax(Q K^T / 8 + key_bias) (dropout) V with Q|K|V = column blocks of qkv [B*S, 3*H*64].
 * Replaces $TF:126-151 (eager_attention_forward) and its autograd backward. */
int b200mm_attention_fwd(const void* qkv, const float* key_bias, void* out, float* lse, int B, int H, int S,
                         float p_drop, unsigned long long seed, void* stream);
int b200mm_attention_bwd(const void* qkv, const float* key_bias, const void* out, const void* dout, const float* lse,
                         void* dqkv, int B, int H, int S, float p_drop, unsigned long long seed, void* stream);
/* int64 attention_mask (1 = token) -> additive fp32 key bias 0 / -inf  ($TF:415-419) */
int b200mm_mask_to_bias(const long long* mask, float* bias, long long n, void* stream);

/* ---- LayerNorm / embeddings ($TF:96-122 Embeddings, :257 sa_layer_norm, :261 output_layer_norm; BERT / XLM-R
 * variants: transformers/models/bert/modeling_bert.py:53-113, xlm_roberta/modeling_xlm_roberta.py:56-159;
 * ViT token assembly: transformers/models/vit/modeling_vit.py ViTEmbeddings.forward) --------------------------- */
int b200mm_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                         int M, int D, float eps, float p_drop, unsigned long long seed, void* stream);
/* x_saved = bf16(word[ids] + pos[pos_ids ? pos_ids[m] : m % S] (+ type_row)); y = dropout(LN(x_saved)).
 * pos_ids: nullable int32 [M] (RoBERTa-style); type_row: nullable fp32 [D] = token_type_embeddings[0]. */
int b200mm_embed_layernorm_fwd(const long long* ids, const float* word, const float* pos, const int* pos_ids,
                               const float* type_row, int S, int vocab, const float* gamma, const float* beta,
                               void* x_saved, void* y, float* mean, float* rstd, int M, int D, float eps,
                               float p_drop, unsigned long long seed, void* stream);
/* addend (nullable bf16 [M,D]) is added to dx only: the residual-stream gradient of a pre-LN (ViT) block. */
int b200mm_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                         const void* addend, void* dx, void* dx2, float* dgamma, float* dbeta, int M, int D,
                         float p_in, unsigned long long seed_in, float p_out, unsigned long long seed_out,
                         void* stream);
int b200mm_embedding_bwd(const void* dx, const long long* ids, const int* pos_ids, long long pos_padding_idx, int S,
                         int vocab, long long padding_idx, float* dword, float* dpos, int M, int D, void* stream);
/* RoBERTa / XLM-R position ids: out[b,s] = cumsum(ids != pad)[b,s] * (ids[b,s] != pad) + pad  (int32 [B*S]) */
int b200mm_position_ids(const long long* ids, long long pad_id, int B, int S, int* out, void* stream);
/* ViT: tokens[B*(P+1), D] = [cls | patch[B*P, D]] + pos[(P+1), D]; backward: dpatch copy, dpos / dcls sums (+=) */
int b200mm_vit_assemble_fwd(const void* patch, const float* cls, const float* pos, void* x, int B, int P, int D,
                            void* stream);
int b200mm_vit_assemble_bwd(const void* dx, void* dpatch, float* dcls, float* dpos, int B, int P, int D,
                            void* stream);
/* pooled token: out[i,:] = dropout(x[i*stride_rows + offset_rows, :])  (h[:, -1, :] + bert_drop,
 * example_scripts/Multimodal_example_task2C.txt:178) and its backward scatter */
int b200mm_gather_rows(const void* x, void* out, int rows, int D, long long stride_rows, long long offset_rows,
                       float p_drop, unsigned long long seed, void* stream);
int b200mm_scatter_rows(const void* dpooled, void* dx, long long M, int D, long long stride_rows,
                        long long offset_rows, float p_drop, unsigned long long seed, void* stream);

/* ---- image tower support: BatchNorm2d (training / eval), pooling, conv lowering ($TV:108-163, :197-206, :266-280)
 * `scratch`: fp32 workspace of at least 18*C + 32 floats, zeroed by the call itself. */
int b200mm_batchnorm_fwd(const void* x, const void* residual, long long M, int C, const float* gamma,
                         const float* beta, float eps, float momentum, int relu, void* out, float* mean_out,
                         float* rstd_out, float* running_mean, float* running_var, float* scratch, void* stream);
/* one-pass variant: col_stats = [sum(C) | sum of squares(C)] accumulated by the producing convolution's epilogue */
int b200mm_batchnorm_fwd_stats(const void* x, const void* residual, long long M, int C, const float* col_stats,
                               const float* gamma, const float* beta, float eps, float momentum, int relu, void* out,
                               float* mean_out, float* rstd_out, float* running_mean, float* running_var,
                               unsigned char* relu_mask, void* stream);
/* relu_mask (nullable, [M, C/8] bytes): 1 bit per element = (pre-ReLU value > 0), all the backward needs of the output */
int b200mm_batchnorm_eval(const void* x, const void* residual, long long M, int C, const float* gamma,
                          const float* beta, const float* running_mean, const float* running_var, float eps, int relu,
                          void* out, void* stream);
/* ReLU mask source, in order of preference: relu_mask (1 bit / element from b200mm_batchnorm_fwd_stats), the saved
 * output `out`, or -- both NULL -- recomputed from x (needs beta; only valid without a residual). */
int b200mm_batchnorm_bwd(const void* dout, const void* out, const void* x, long long M, int C, const float* mean,
                         const float* rstd, const float* gamma, const float* beta, const unsigned char* relu_mask,
                         int relu, void* dx, void* dz_out, float* dgamma, float* dbeta, float* scratch, void* stream);
int b200mm_maxpool3x3s2_fwd(const void* x, int N, int H, int W, int C, void* out, void* argmax, void* stream);
/* Stem tail in one pass: out = maxpool3x3s2(relu(BN_train(x))) straight from the convolution output x [N,H,W,C] whose
 * column statistics are in col_stats; the normalised activation is never materialised.  Bit-identical to
 * b200mm_batchnorm_fwd_stats + b200mm_maxpool3x3s2_fwd.  Replaces bn1 / relu / maxpool of
 * torchvision/models/resnet.py:268-271. */
int b200mm_bn_relu_maxpool_fwd(const void* x, int N, int H, int W, int C, const float* col_stats, const float* gamma,
                               const float* beta, float eps, float momentum, void* out, void* argmax, float* mean_out,
                               float* rstd_out, float* running_mean, float* running_var, void* stream);
int b200mm_maxpool3x3s2_bwd(const void* dout, const void* argmax, int N, int H, int W, int C, void* dx, void* stream);
int b200mm_avgpool_fwd(const void* x, int N, int HW, int C, void* out, void* stream);
int b200mm_avgpool_bwd(const void* dout, int N, int HW, int C, void* dx, void* stream);
int b200mm_im2col_nhwc(const void* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad, void* cols,
                       void* stream);
int b200mm_col2im_nhwc(const void* dcols, const void* addend, int N, int H, int W, int C, int KH, int KW, int stride,
                       int pad, void* dx, void* stream);
int b200mm_im2col_nchw_f32(const float* img, int N, int Cin, int H, int W, int KH, int KW, int stride, int pad,
                           int Kp, void* cols, void* stream);
/* ResNet stem (torchvision/models/resnet.py:197 conv1, 7x7 / 2 / pad 3, 3 -> 64) computed from the fp32 NCHW image
 * without an im2col matrix: forward (+ optional BatchNorm column statistics, accumulated into col_stats[128]) and the
 * weight gradient (dw fp32 [64,152] accumulated).  B200MM_ERR_BAD_ARG for any other stem shape (Cin != 3, Cout != 64,
 * Kp != 152, W > 226 or Wo > 128): the caller then lowers through b200mm_im2col_nchw_f32 + b200mm_gemm_bf16. */
int b200mm_stem_conv_fwd(const float* img, int N, int Cin, int H, int W, const void* w, int Cout, int Kp, void* out,
                         float* col_stats, void* stream);
int b200mm_stem_conv_wgrad(const float* img, int N, int Cin, int H, int W, const void* dy, int Cout, int Kp,
                           float* dw, void* stream);
int b200mm_subsample_nhwc(const void* x, int N, int H, int W, int C, int stride, void* out, void* stream);
int b200mm_upsample_add_nhwc(const void* dsub, const void* addend, int N, int H, int W, int C, int stride, void* dx,
                             void* stream);

/* ---- image preprocessing: uint8 HWC -> Resize(shorter side, antialiased bilinear) -> CenterCrop -> /255 -> Normalize ->
 * fp32 NCHW, fused (example_scripts/Multimodal_example_task2C.txt:37-41 transforms).  `images`: device array of n device
 * pointers; mean3/std3: HOST arrays of 3 floats. */
int b200mm_preprocess_u8(const void* images, const int* heights, const int* widths, int n, int resize, int crop,
                         const float* mean3, const float* std3, float* out, void* stream);
/* Same transform over a PACKED batch (one byte buffer + device tables of offsets / heights / widths: two H2D copies per
 * batch whatever its size); square = 1 selects Resize((crop, crop)) of the HEAD script
 * (example_scripts/Multimodal_example_task2C.py:222-235), flip = per-image RandomHorizontalFlip flags or NULL. */
int b200mm_preprocess_u8_packed(const void* packed, const long long* offsets, const int* heights, const int* widths,
                                const void* flip, int n, int resize, int crop, int square, const float* mean3,
                                const float* std3, float* out, void* stream);
/* Same contract, PILLOW-EXACT: the resize is Pillow's 8-bit two-pass bilinear (22-bit fixed-point coefficients, uint8
 * between the passes; src/libImaging/Resample.c restated in csrc/resample_math.cuh), ToTensor / Normalize with torch's
 * roundings -- the output equals the tensor the reference's Dataset builds from the PIL image bit for bit
 * (example_scripts/Multimodal_example_task2C.txt:37-41, :50).  Sides that shrink by more than 31x are not supported. */
int b200mm_preprocess_u8_packed_pil(const void* packed, const long long* offsets, const int* heights, const int* widths,
                                    const void* flip, int n, int resize, int crop, int square, const float* mean3,
                                    const float* std3, float* out, void* stream);
/* The Pillow-exact Resize / CenterCrop / flip as uint8: out [n, crop, crop, 3] -- input of b200mm_augment_pil. */
int b200mm_preprocess_u8_packed_pil_u8(const void* packed, const long long* offsets, const int* heights, const int* widths,
                                       const void* flip, int n, int resize, int crop, int square, void* out, void* stream);
/* ColorJitter + RandomRotation + ToTensor + Normalize of example_scripts/Multimodal_example_task2C.py:224-235 with Pillow's
 * own uint8 arithmetic (ImageEnhance blends, convert("L") / ("HSV"), Image.rotate's 16.16 fixed-point affine; restated in
 * csrc/augment_pil_math.cuh): img_u8 [n, H, W, 3]; order int [n] (2 bits per operator, first applied in the low bits);
 * alpha fp32 [n, 3] = brightness / contrast / saturation factors; hue int [n] = uint8(hue_factor * 255); affine int [n, 6]
 * = the rotation matrix in 16.16 fixed point as Pillow's affine_fixed prepares it; sums: n x uint64 scratch; out fp32
 * [n, 3, H, W].  The output equals the tensor the script's Dataset builds for the same draws, bit for bit. */
int b200mm_augment_pil(const void* img_u8, const int* order, const float* alpha, const int* hue, const int* affine, int n,
                       int H, int W, const float* mean3, const float* std3, void* sums, float* out, void* stream);
/* Batch already at network resolution: [n, H, W, 3] uint8 -> ToTensor -> Normalize -> fp32 NCHW (W % 4 == 0). */
int b200mm_u8_normalize_nchw(const void* src, const void* flip, int n, int H, int W, const float* mean3,
                             const float* std3, float* out, void* stream);
/* Train-time augmentations of the HEAD script (example_scripts/Multimodal_example_task2C.py:224-233): ColorJitter (four
 * operators in a per-image random order) + RandomRotation (nearest, zero fill) + Normalize over a batch already resized /
 * flipped / scaled to [0, 1] (the two entry points above with mean 0, std 1).  img01, out: [n, 3, H, W] fp32 (out != img01);
 * order [n] int: 2 bits per operator, first applied in the low bits (0 brightness, 1 contrast, 2 saturation, 3 hue);
 * params [n, 8] fp32: brightness / contrast / saturation factors, hue shift, inverse rotation matrix m00 m01 m10 m11;
 * gray_mean [9 n] fp32: scratch; entries [0, n) hold each image's contrast mean afterwards.  Semantics: torchvision's
 * float-tensor path. */
int b200mm_augment_jitter_rotate(const float* img01, const int* order, const float* params, int n, int H, int W,
                                 const float* mean3, const float* std3, float* gray_mean, float* out, void* stream);

/* ---- JPEG decode, split for the GPU (the decode inside the reference's Dataset: `Image.open(path).convert("RGB")`,
 * example_scripts/Multimodal_example_task2C.txt:50, Multimodal_example_task2C.py:270).  HOST functions (no CUDA call; meant
 * for loader workers): marker parsing + Huffman decoding of 8-bit baseline / progressive files, grayscale or YCbCr with
 * 4:4:4 / 4:2:2 / 4:2:0 sampling.  Return 0, B200MM_JPEG_UNSUPPORTED (-10: arithmetic coding, 12-bit, CMYK, RGB-coded,
 * other sampling factors -- decode those with the caller's loader) or B200MM_JPEG_CORRUPT (-11, incl. truncated files).
 * info: int[32] = width, height, components, progressive, hs, vs, blocks per row [3], block rows [3] (both padded to whole
 * MCUs), real component width [3], height [3], first coefficient of each component [3], coefficients in total, restart
 * interval.  coefs: info[21] int16, 64 per block in natural order; qtabs: [3][64] uint16, natural order. */
int b200mm_jpeg_parse(const void* data, long long len, int* info);
int b200mm_jpeg_entropy_decode(const void* data, long long len, short* coefs, unsigned short* qtabs, int* info);
/* DEVICE: dequantisation + inverse DCT + chroma up-sampling + YCbCr -> RGB of a whole batch in two launches, bit-identical to
 * libjpeg-turbo's default path (JDCT_ISLOW, fancy up-sampling), i.e. to Pillow's pixels.  coefs: the batch's coefficients;
 * qtabs [n][3][64]; table: int64 [n][32] = width, height, components (0: skip, the caller supplies the pixels), hs, vs,
 * blocks per row [3], block rows [3], component width [3], height [3], first coefficient in `coefs` [3] (int16 elements,
 * multiples of 8), first byte of each component plane in `planes` [3] (multiples of 8), first byte of the image in `out`,
 * blocks of the image in total.  out: packed uint8 RGB, image i = [height][width][3] -- the layout
 * b200mm_preprocess_u8_packed reads.  max_blocks / max_w / max_h: maxima over the batch (grid sizing). */
int b200mm_jpeg_reconstruct(const short* coefs, const unsigned short* qtabs, const long long* table, int n, int max_blocks,
                            int max_w, int max_h, void* planes, void* out, void* stream);
/* The same pair with the coefficients delivered SPARSE (a typical file keeps 5-15 % of them; 3 B per non-zero + 4 B per
 * block cross PCIe instead of 128 B per block).  HOST: scratch = info[21] int16 work space; block_off int32 [blocks + 1],
 * idx uint8 [nnz] (position inside the block, natural order), val int16 [nnz]; capacity = entries idx / val can hold
 * (info[21] always suffices); *nnz receives the count.  DEVICE: sp_off = the batch's block_off arrays concatenated with
 * batch-wide entry offsets ([blocks of the batch + 1]); table columns 17..19 = first BLOCK of each component in the
 * batch's block numbering. */
int b200mm_jpeg_entropy_decode_sparse(const void* data, long long len, short* scratch, int* block_off, unsigned char* idx,
                                      short* val, long long capacity, unsigned short* qtabs, int* info, long long* nnz);
int b200mm_jpeg_reconstruct_sparse(const int* sp_off, const unsigned char* sp_idx, const short* sp_val,
                                   const unsigned short* qtabs, const long long* table, int n, int max_blocks, int max_w,
                                   int max_h, void* planes, void* out, void* stream);

/* ---- head + loss, optimizer --------------------------------------------------------------------------------------
 * output layer fused with the loss: example_scripts/Multimodal_example_task2C.txt:195 (output_fc) + :214, :248
 * (nn.CrossEntropyLoss); loss_kind 1 = torchvision.ops.sigmoid_focal_loss (Multimodal_example_task2C.py:167);
 * loss_kind 2 = backward only, from an external dL/dlogits. */
int b200mm_head_loss(const void* feat, const float* W, const float* bias, const long long* labels, int B, int F,
                     int C, int loss_kind, float alpha, float gamma, int train, const float* dlogits_in,
                     float* logits, float* loss_sum, int* correct, void* dfeat, float* dW, float* dbias,
                     void* stream);
int b200mm_colsum_bf16(const void* x, long long ld, int M, int N, float* out, void* stream);
/* ---- HEAD-script head (example_scripts/Multimodal_example_task2C.py:476-499 ConcatAttention3, :571-574 fine_tune,
 * :599-601 text_fc, :641-643 output_fc, :167 sigmoid_focal_loss): small batch-dimension kernels ------------------- */
/* BatchNorm1d (+ReLU) over x[B,C] (bf16, or fp32 when x_f32 != 0); train: batch statistics (saved in mean/rstd,
 * folded into running_*); out / dx are bf16. */
int b200mm_bn1d_fwd(const void* x, int x_f32, long long ldx, int B, int C, const float* gamma, const float* beta, float eps,
                    float momentum, int relu, int train, void* out, long long ldo, float* mean, float* rstd,
                    float* running_mean, float* running_var, void* stream);
int b200mm_bn1d_bwd(const void* dout, long long ldd, const void* out, long long ldo, const void* x, int x_f32,
                    long long ldx, int B, int C, const float* mean, const float* rstd, const float* gamma, int relu, void* dx,
                    long long lddx, float* dgamma, float* dbeta, void* stream);
/* y = softmax(a, dim=1) * x (weights w kept in fp32); backward: da, and dx_direct = dy * w */
int b200mm_softmax_gate_fwd(const void* a, const void* x, int B, int C, float* w, void* y, void* stream);
int b200mm_softmax_gate_bwd(const void* dy, const float* w, const void* x, int B, int C, void* da, void* dx_direct,
                            void* stream);
int b200mm_relu_bwd(const void* dy, const void* y, long long n, void* dx, void* stream);
/* Linear(F,1) + BatchNorm1d(1) + sigmoid focal loss (+ backward), one CTA, B <= 4096; loss / correct accumulate. */
int b200mm_head_bn_focal(const void* feat, const float* W, const float* bias, const float* bn_gamma,
                         const float* bn_beta, float* running_mean, float* running_var, const long long* labels, int B,
                         int F, float eps, float momentum, float alpha, float gamma, int train, int bn_train,
                         const float* dlogits_in, float* logits, float* loss, int* correct, void* dfeat, float* dW,
                         float* dbias, float* dg, float* dbeta, void* stream);

/* optim.Adam(...).step() (example_scripts/Multimodal_example_task2C.txt:217, :249) + clip_grad_norm_
 * (Multimodal_example_task2C.py:713-715) over a flat fp32 parameter range, refreshing the bf16 shadow */
int b200mm_sumsq_f32(const float* g, long long n, float* out, void* stream);
int b200mm_adam_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int step, const float* gradsq,
                     float max_norm, float grad_scale, void* stream);
int b200mm_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream);
/* Data-parallel gradient payload in bf16 ("bf16 payload -> fp32 master update"): pack a gradient range
 * pre-scaled by 1 / world size, and the optimizer / clip-norm variants that read the all-reduced bf16 buffer. */
int b200mm_scale_cast_f32_to_bf16(const float* x, void* y, long long n, float scale, void* stream);
int b200mm_sumsq_bf16(const void* g_bf16, long long n, float* out, void* stream);
int b200mm_adam_step_g16(float* p, const void* g_bf16, float* m, float* v, void* shadow_bf16, long long n, float lr,
                         float beta1, float beta2, float eps, float weight_decay, int step, const float* gradsq,
                         float max_norm, float grad_scale, void* stream);

/* Attention with the dropout keep bits handed from the forward to the backward (one-tile sequences, S <= 128):
 * drop_mask = [B*H][4][128] uint32 device buffer.  Same op as b200mm_attention_fwd / _bwd
 * (transformers/models/distilbert/modeling_distilbert.py:126-151 and its autograd backward). */
int b200mm_attention_fwd_mask(const void* qkv, const float* key_bias, void* out, float* lse, int B, int H, int S,
                              float p_drop, unsigned long long seed, void* drop_mask, void* stream);
int b200mm_attention_bwd_mask(const void* qkv, const float* key_bias, const void* out, const void* dout,
                              const float* lse, void* dqkv, int B, int H, int S, float p_drop,
                              unsigned long long seed, const void* drop_mask, void* stream);

/* Feature-extraction path (baselines/extract_feat.py:52-67): ConvNeXt's depthwise 7x7 convolution
 * (torchvision/models/convnext.py CNBlock; NHWC bf16, wt = [49][C] tap-major) and BertPooler's tanh. */
int b200mm_dwconv7x7_nhwc(const void* x, const void* wt, const float* bias, void* y, int N, int H, int W, int C,
                          void* stream);
int b200mm_tanh_f32(float* x, long long n, void* stream);

/* CUDA-graph support for the whole train step (small-batch regime: the reference's batch 16 / 8 is launch-bound).
 * A captured graph freezes by-value launch parameters, so the three per-step host values move to device memory:
 * the dropout seed salt (added to every dropout seed by the kernels), the Adam step count and the learning rate.
 * Reference loop being captured: example_scripts/Multimodal_example_task2C.txt:204-217. */
int b200mm_set_step_salt_ptr(const unsigned long long* salt_dev);
int b200mm_adam_step_dyn(float* p, const void* g, int g_is_bf16, float* m, float* v, void* shadow_bf16, long long n,
                         const float* lr_dev, float beta1, float beta2, float eps, float weight_decay,
                         const int* step_dev, const float* gradsq, float max_norm, float grad_scale, void* stream);
int b200mm_step_advance(unsigned long long* salt_dev, int* adam_step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MM_H_ */
