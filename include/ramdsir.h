/*
 * ramdsir.h -- C ABI of the B200-native Random Amplitude Mixup (RAM) library (libramdsir.so).
 *
 * This is the drop-in boundary for the one hot path of zzzqzhou/RAM-DSIR that this repo
 * re-implements.  Every entry point names the reference interface it replaces
 * (paths relative to the reference checkout, code/dataset/fundus.py == code/dataset/prostate.py):
 *
 *   ramdsir_bank_build_f32   <- extract_amp_spectrum(img_np)            fundus.py:13-19
 *                               (only the low band the blend reads: fftshift(amp)[:, h1:h2, w1:w2],
 *                                fundus.py:22-33)
 *   ramdsir_mix_f32          <- source_to_target_freq(src, amp_trg, L)  fundus.py:41-61
 *                               + low_freq_mutate_np                    fundus.py:21-39
 *                               + caller epilogue clip/normalise        fundus.py:215,222-223 / prostate.py:188
 *   ramdsir_mix_host_f32     <- the same, as called from the data loader with HOST arrays
 *                               (Fundus_Multi.__getitem__ fundus.py:197-225, Prostate_Multi prostate.py:177-202)
 *   ramdsir_band_halfwidth   <- b = floor(min(h, w) * L)                fundus.py:26
 *
 * Conventions
 *   - plain C, no C++/torch types; all *_dev pointers are BORROWED device pointers on the current
 *     CUDA device, densely packed NCHW float32 (16-byte aligned unless the shape runs on the generic kernels); `stream` is a cudaStream_t cast to
 *     void* (NULL = legacy default stream).  Calls are asynchronous and stream-ordered.
 *   - the library owns nothing persistent except per-(device, H, W, b) twiddle tables built lazily
 *     on first use (a few KB each) and, for the *_host entry, an explicit pipeline handle.
 *   - every function returns RAMDSIR_OK (0) or an error code; ramdsir_last_error() gives the
 *     thread-local message.  There is no CPU fallback and no cuFFT: an unsupported shape is an error.
 *   - amplitude bank layout: bank[n][c][kh + b][kw + b] = |FFT2(img n, channel c)[kh mod H, kw mod W]|,
 *     kh, kw in [-b, b]  (numpy's unnormalised forward transform), i.e. exactly
 *     np.fft.fftshift(extract_amp_spectrum(img))[:, cH-b:cH+b+1, cW-b:cW+b+1].
 *   - partner selection and ratio are drawn on the host with the reference's own RNG calls and passed
 *     in: partner_idx[i] indexes the bank, ratio = ratio_k[i] / ratio_den weights the SOURCE amplitude
 *     (fundus.py:35,37; the reference uses ratio_k in 1..10, ratio_den = 10).  An out-of-range
 *     partner index poisons that image's output with NaN instead of reading out of bounds.
 *   - epilogue: out = clamp(v, clip_lo, clip_hi) * scale + shift.  fundus: (0, 255, 1/127.5, -1);
 *     prostate: (-1, 1, 1, 0); "no epilogue": (-INF, INF, 1, 0).
 */
#ifndef RAMDSIR_H_
#define RAMDSIR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAMDSIR_OK 0
#define RAMDSIR_ERR_INVALID 1      /* bad argument (null pointer, non-positive size, misalignment) */
#define RAMDSIR_ERR_UNSUPPORTED 2  /* shape / band not supported by any CUDA path */
#define RAMDSIR_ERR_CUDA 3         /* a CUDA runtime call failed; see ramdsir_last_error() */
#define RAMDSIR_ERR_WORKSPACE 4    /* workspace too small; see ramdsir_workspace_bytes() */

/* kernel families, as reported by ramdsir_path() */
#define RAMDSIR_PATH_NONE 0
#define RAMDSIR_PATH_FUSED 1      /* one CTA per plane, band buffers resident in shared memory */
#define RAMDSIR_PATH_MULTIPASS 2  /* large images: band buffers in an L2-resident workspace */
#define RAMDSIR_PATH_GENERIC 3    /* any H, W: direct band-pruned DFT kernels */

int ramdsir_version(void);
const char* ramdsir_last_error(void);

/* b = floor(min(H, W) * L) evaluated in float64, as fundus.py:26 does. */
int ramdsir_band_halfwidth(int H, int W, double L);

/* Which kernel family serves this shape (RAMDSIR_PATH_*). */
int ramdsir_path(int C, int H, int W, int b);

/*
 * Optional warm-up: builds the per-(device, H, W, b) twiddle tables now instead of on the first call.  Table
 * construction uploads a few KB on a private stream and synchronises THAT stream only (never the device, never the
 * caller's streams); call this once per shape outside a latency-critical region if even that matters.
 */
int ramdsir_prepare(int C, int H, int W, int b);

/*
 * Per-call flags (the `flags` argument of the *_ex / dual entry points).  There is no process-global state.
 *
 * Overlap of consecutive launches.  By default the static fused kernels decide by themselves whether a launch may
 * start on SMs that the previous RAM launch in the same stream has already vacated (programmatic dependent launch):
 * the library knows every buffer of both launches and overlaps them only if they touch disjoint memory (no
 * read-after-write, write-after-read or write-after-write between them).  Either way the later launch does not
 * COMPLETE before the earlier one (griddepcontrol.wait at its end), so ordering against copies, events and every other
 * kernel in the stream is the plain stream order.
 *   RAMDSIR_FLAG_INDEPENDENT  skip the check (the caller asserts the buffers are disjoint)
 *   RAMDSIR_FLAG_SERIAL       never overlap: launch strictly after the previous kernel has finished
 * Kernel-family selection, for cross-checking two independent CUDA implementations in tests:
 *   RAMDSIR_FLAG_PATH_GENERIC   use the generic (any-shape, direct band-pruned DFT) kernels
 *   RAMDSIR_FLAG_RUNTIME_SHAPE  on the fused path, use the runtime-shape kernel even where a static one exists
 * (Setting the environment variable RAMDSIR_NO_AUTO_OVERLAP=1 before loading the library disables the automatic
 * decision, leaving only RAMDSIR_FLAG_INDEPENDENT.)
 */
#define RAMDSIR_FLAG_INDEPENDENT 1u
#define RAMDSIR_FLAG_SERIAL 2u
#define RAMDSIR_FLAG_PATH_GENERIC 4u
#define RAMDSIR_FLAG_RUNTIME_SHAPE 8u

/* Scratch bytes needed to process `n_images` images of this shape in one call (0 for the fused path). */
size_t ramdsir_workspace_bytes(int n_images, int C, int H, int W, int b);
size_t ramdsir_workspace_bytes_ex(int n_images, int C, int H, int W, int b, unsigned flags);

/* K3: amplitude bank.  imgs_dev [N, C, H, W] -> bank_dev [N, C, 2b+1, 2b+1]. */
int ramdsir_bank_build_f32(const float* imgs_dev, float* bank_dev, int N, int C, int H, int W, int b,
                           void* workspace_dev, size_t workspace_bytes, void* stream);
int ramdsir_bank_build_f32_ex(const float* imgs_dev, float* bank_dev, int N, int C, int H, int W, int b,
                              void* workspace_dev, size_t workspace_bytes, void* stream, unsigned flags);

/* K1/K2: RAM.  src_dev [B, C, H, W] -> out_dev [B, C, H, W]; out_dev must not overlap src_dev (the source is read twice). */
int ramdsir_mix_f32(const float* src_dev, const float* bank_dev, const int64_t* partner_idx_dev,
                    const int32_t* ratio_k_dev, float* out_dev, int B, int C, int H, int W, int b,
                    int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift,
                    void* workspace_dev, size_t workspace_bytes, void* stream);

/* Same as ramdsir_mix_f32 with RAMDSIR_FLAG_* (see above). */
int ramdsir_mix_f32_ex(const float* src_dev, const float* bank_dev, const int64_t* partner_idx_dev,
                       const int32_t* ratio_k_dev, float* out_dev, int B, int C, int H, int W, int b,
                       int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift,
                       void* workspace_dev, size_t workspace_bytes, void* stream, unsigned flags);

/*
 * Dual output (what the train step consumes, train.py:244): besides out = clamp(RAM(src)) * scale + shift, also
 * write the normalised CLEAN batch clean_out = src * scale + shift (fundus.py:217-218: img /= 127.5; img -= 1.0,
 * reproduced to 1 ulp).  On the static fused kernels it is produced from the same read of src (no extra pass over
 * the input); other paths run one extra streaming kernel.  clean_out_dev may be NULL (= ramdsir_mix_f32_ex).
 */
int ramdsir_mix_dual_f32(const float* src_dev, const float* bank_dev, const int64_t* partner_idx_dev,
                         const int32_t* ratio_k_dev, float* out_dev, float* clean_out_dev, int B, int C, int H, int W,
                         int b, int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift,
                         void* workspace_dev, size_t workspace_bytes, void* stream, unsigned flags);

/*
 * uint8 sources.  The reference's loaders hand RAM pixels that are uint8-valued (PIL image -> np.array ->
 * astype(float32), fundus.py:205-214; the bank sources likewise, fundus.py:210-214).  These entry points read the
 * uint8 planes [.., C, H, W] directly (1 byte instead of 4 per pixel over PCIe and HBM); the arithmetic is the
 * float32 path's on float(src), so results are BIT-IDENTICAL to ramdsir_mix_dual_f32 / ramdsir_bank_build_f32 on
 * the converted input.  Available on the static fused kernels only (ramdsir_u8_supported() != 0: 256x256 with
 * b = 25, 384x384 with b = 38); other shapes return RAMDSIR_ERR_UNSUPPORTED (convert on the device and use the
 * f32 entry points).  src must be 4-byte aligned, out / clean_out 16-byte aligned; clean_out_dev may be NULL.
 */
int ramdsir_u8_supported(int C, int H, int W, int b);
int ramdsir_bank_build_u8(const uint8_t* imgs_dev, float* bank_dev, int N, int C, int H, int W, int b, void* stream);
int ramdsir_mix_dual_u8(const uint8_t* src_dev, const float* bank_dev, const int64_t* partner_idx_dev,
                        const int32_t* ratio_k_dev, float* out_dev, float* clean_out_dev, int B, int C, int H, int W,
                        int b, int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale, float shift,
                        void* stream, unsigned flags);

/*
 * Strict per-sample drop-ins (correctness shims for code that calls the reference functions directly).
 *   ramdsir_amp_spectrum_f32 <- extract_amp_spectrum: FULL amplitude spectrum [N, C, H, W], un-shifted
 *                               (fundus.py:13-19).  Direct DFT, any H and W; workspace from
 *                               ramdsir_amp_spectrum_workspace_bytes().
 *   ramdsir_band_blend_f32   <- low_freq_mutate_np with the ratio numerator supplied by the caller
 *                               (fundus.py:21-39): out = amp_src, band blended with amp_trg.
 */
size_t ramdsir_amp_spectrum_workspace_bytes(int N, int C, int H, int W);
int ramdsir_amp_spectrum_f32(const float* imgs_dev, float* amp_dev, int N, int C, int H, int W,
                             void* workspace_dev, size_t workspace_bytes, void* stream);
int ramdsir_band_blend_f32(const float* amp_src_dev, const float* amp_trg_dev, float* out_dev, int C, int H, int W,
                           int b, int ratio_k, int ratio_den, void* stream);

/*
 * Host-buffer entry point (what a data-loader process calls).  A pipeline handle owns device
 * staging buffers and three streams; ramdsir_mix_host_f32 splits the batch into chunks and
 * overlaps host->device copy, the RAM kernel and device->host copy.  src_host / out_host should be
 * pinned (cudaHostAlloc / torch pin_memory) for the copies to be asynchronous; pageable memory
 * works but serialises.  The call returns after out_host is complete.
 */
typedef struct ramdsir_pipe ramdsir_pipe_t;
int ramdsir_pipe_create(ramdsir_pipe_t** pipe, int max_chunk_images, int C, int H, int W, int b);
void ramdsir_pipe_destroy(ramdsir_pipe_t* pipe);
int ramdsir_mix_host_f32(ramdsir_pipe_t* pipe, const float* src_host, const float* bank_dev,
                         const int64_t* partner_idx_host, const int32_t* ratio_k_host, float* out_host,
                         int B, int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale,
                         float shift);
/* Same with a uint8 source batch [B, C, H, W] (what a fundus loader holds before astype(float32), fundus.py:212):
 * a quarter of the host->device bytes; widened to float32 on the device.  Output is float32 as before. */
int ramdsir_mix_host_u8(ramdsir_pipe_t* pipe, const unsigned char* src_host, const float* bank_dev,
                        const int64_t* partner_idx_host, const int32_t* ratio_k_host, float* out_host,
                        int B, int n_bank, int ratio_den, float clip_lo, float clip_hi, float scale,
                        float shift);
/*
 * Device-side partner / ratio selection with host-identical MT19937 streams (what Fundus_Multi.__getitem__ draws per
 * sample: np.random.choice(domain_list, 1)[0], np.random.choice(partner_list), random.randint(1, 10) --
 * fundus.py:205,208,35, prostate.py:182-183 -- from the per-worker streams seeded by seed_worker, train.py:80-83).
 * A handle owns one (numpy, python) MT19937 state pair per worker stream ON THE DEVICE, seeded exactly like
 * np.random.seed(s) / random.seed(s).  Stream s draws partners among candidates[s][0 .. n_candidates[s]) (domain
 * indices; the caller removes the test domain and, for is_out_domain, the stream's own domain) and, inside the chosen
 * domain d, an image in [0, domain_sizes[d]).  ramdsir_rng_draw produces, for every stream, the next n_per_stream
 * samples' selections: partner_idx_dev[s * n + i] = row of the concatenated bank (domain offset + image index), and
 * ratio_k_dev[s * n + i] in [ratio_lo, ratio_hi] -- bit-identical to calling the reference's draws in a loop on the
 * host, including the number of 32-bit outputs each rejection loop consumes.  `candidates` is [n_streams][8].
 * ramdsir_rng_seed_host / ramdsir_rng_get_state expose the raw states in numpy's get_state() layout (624 words +
 * position; ramdsir_rng_state_words() == 625) so the streams can be handed back and forth with the host generators.
 */
typedef struct ramdsir_rng ramdsir_rng_t;
int ramdsir_rng_state_words(void);
int ramdsir_rng_seed_host(uint32_t worker_seed, uint32_t* numpy_state, uint32_t* python_state);
int ramdsir_rng_create(ramdsir_rng_t** rng, int n_streams, const uint32_t* worker_seeds, int n_domains,
                       const int32_t* domain_sizes, const int32_t* n_candidates, const int32_t* candidates);
void ramdsir_rng_destroy(ramdsir_rng_t* rng);
int ramdsir_rng_draw(ramdsir_rng_t* rng, int n_per_stream, int64_t* partner_idx_dev, int32_t* ratio_k_dev, int ratio_lo,
                     int ratio_hi, void* stream);
int ramdsir_rng_get_state(ramdsir_rng_t* rng, int stream_index, uint32_t* numpy_state, uint32_t* python_state);

/*
 * Device-side fundus mask encoding (fundus.py:227-239 + to_multilabel, transform.py:10-14): 8-bit mask [B, H, W]
 * (grey > 200 background, 51..200 disc rim, <= 50 cup) -> float32 multilabel [B, 2, H, W]: channel 0 = cup,
 * channel 1 = disc (cup included).  Byte work: bit-exact.
 */
int ramdsir_encode_fundus_mask_u8(const uint8_t* mask_dev, float* out_dev, int B, int H, int W, void* stream);

/*
 * GPU versions of the loader's geometric transforms on 8-bit images, bit-exact with Pillow (which the reference calls):
 *   ramdsir_resize_u8        <- img.resize(size, BILINEAR) / mask.resize(size, NEAREST)   dataset/transform.py:163-177,
 *                               193-196; partner resize dataset/fundus.py:209
 *   ramdsir_crop_pad_chw_u8  <- RandomCrop's right/bottom padding + crop                  dataset/transform.py:25-38
 *                               (+ the HWC -> CHW transpose the RAM block does next,      dataset/fundus.py:211-212)
 * Images are interleaved HWC uint8 on the device, as np.array(PIL image) lays them out.  `filter` uses Pillow's numbering.
 * Bilinear is Pillow's two-pass fixed-point convolution (horizontal first, 8-bit intermediate, 22-bit coefficients),
 * nearest its running-sum index table; the tables are built on the host exactly as Pillow builds them
 * (ramdsir_resize_tables_host exposes them: bilinear -> bounds [out][2] = (first, count), kk [out][ksize];
 * nearest -> bounds [out] = source index).  x0 / y0 of the crop may be negative or run past the image: pixels outside
 * are `fill` (the reference pads images with 0 and masks with 255).
 */
#define RAMDSIR_FILTER_NEAREST 0
#define RAMDSIR_FILTER_BILINEAR 2
size_t ramdsir_resize_workspace_bytes(int Hin, int Win, int C, int Hout, int Wout);
int ramdsir_resize_u8(const uint8_t* src_dev, int Hin, int Win, int C, uint8_t* dst_dev, int Hout, int Wout, int filter,
                      void* workspace_dev, size_t workspace_bytes, void* stream);
int ramdsir_crop_pad_chw_u8(const uint8_t* src_hwc_dev, int H, int W, int C, int x0, int y0, int out_h, int out_w,
                            int fill, uint8_t* dst_chw_dev, void* stream);
int ramdsir_resize_tables_host(int in_size, int out_size, int filter, int32_t* bounds, int32_t* kk, int* ksize);

/* number of kernel launches the library has issued on this process so far (bench bookkeeping) */
uint64_t ramdsir_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RAMDSIR_H_ */
