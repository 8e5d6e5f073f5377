// Device-side partner / ratio selection with host-identical MT19937 streams, and device-side
// fundus mask encoding (8f rank 4, second half).
//
// Reference semantics (paths relative to /root/reference/code):
//   partner domain   np.random.choice(domain_list, 1)[0]     dataset/fundus.py:205, dataset/prostate.py:182
//   partner image    np.random.choice(lines_or_files)        dataset/fundus.py:208, dataset/prostate.py:183
//   blend ratio      random.randint(1, 10) / 10              dataset/fundus.py:35
//   stream seeding   np.random.seed(s); random.seed(s)       train.py:80-83 (per DataLoader worker)
// numpy's legacy global RandomState and CPython's `random` are both MT19937; they differ in seeding and in how a
// bounded integer is taken from the 32-bit outputs:
//   numpy   seed(int)          = init_genrand(s)
//           randint(0, n)      = masked rejection: draw u32 & mask (mask = 2^ceil(log2 n) - 1) until <= n - 1;
//                                n == 1 consumes nothing                  (numpy/random/src/distributions: bounded_masked_uint32)
//   CPython seed(int)          = init_by_array({s})   for 0 <= s < 2^32
//           randint(a, b)      = a + _randbelow(b - a + 1): k = bit_length(width); draw u32 >> (32 - k) until < width
// Every worker stream is one (numpy state, python state) pair; one warp per stream regenerates the 624-word state in
// three data-parallel phases and lane 0 consumes the tempered outputs in order, because each draw's bound depends on
// the previous draw's result (the image bound is the size of the domain just chosen).
#include <cstdint>
#include <cstring>
#include <vector>
#include "ram_internal.h"

namespace ramdsir {

constexpr int MT_N = 624, MT_M = 397;
constexpr int RNG_STATE_WORDS = MT_N + 1;      // 624 state words + the output position
constexpr int RNG_MAX_CAND = 8;

// ---- host: seeding exactly as numpy / CPython do ------------------------------------------------------------------
static void mt_init_genrand(uint32_t* mt, uint32_t s) {
  mt[0] = s;
  for (int i = 1; i < MT_N; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  mt[MT_N] = MT_N;                               // position: state exhausted, the first draw twists
}

static void mt_init_by_array(uint32_t* mt, const uint32_t* key, int len) {
  mt_init_genrand(mt, 19650218u);
  int i = 1, j = 0;
  for (int k = (MT_N > len ? MT_N : len); k; --k) {
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    if (++i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
    if (++j >= len) j = 0;
  }
  for (int k = MT_N - 1; k; --k) {
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    if (++i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
  }
  mt[0] = 0x80000000u;
  mt[MT_N] = MT_N;
}

// ---- device ---------------------------------------------------------------------------------------------------------
struct MtWarp {
  uint32_t* cur;      // shared: 624 tempered outputs of the current block
  uint32_t* a;        // shared: raw state (old)
  uint32_t* b;        // shared: raw state (new)
  int pos;            // next output (uniform across the warp)
};

__device__ __forceinline__ uint32_t mt_mix(uint32_t u, uint32_t v) {
  const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// One twist of the whole state by a warp: new[k] = old/new[k + 397 mod 624] ^ mix(old[k], old[k + 1]).
__device__ void mt_twist(MtWarp& g, int lane) {
  const uint32_t* o = g.a;
  uint32_t* n = g.b;
  for (int k = lane; k < MT_N - MT_M; k += 32) n[k] = o[k + MT_M] ^ mt_mix(o[k], o[k + 1]);                     // 0..226
  __syncwarp();
  for (int k = MT_N - MT_M + lane; k < 2 * (MT_N - MT_M); k += 32) n[k] = n[k - (MT_N - MT_M)] ^ mt_mix(o[k], o[k + 1]);   // 227..453
  __syncwarp();
  for (int k = 2 * (MT_N - MT_M) + lane; k < MT_N - 1; k += 32) n[k] = n[k - (MT_N - MT_M)] ^ mt_mix(o[k], o[k + 1]);      // 454..622
  __syncwarp();
  if (lane == 0) n[MT_N - 1] = n[MT_M - 1] ^ mt_mix(o[MT_N - 1], n[0]);
  __syncwarp();
  for (int k = lane; k < MT_N; k += 32) {
    uint32_t y = n[k];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    g.cur[k] = y;
  }
  __syncwarp();
  uint32_t* t = g.a; g.a = g.b; g.b = t;       // the new raw state is the next twist's input
  g.pos = 0;
}

// lane 0 only, between twists
__device__ __forceinline__ bool mt_have(const MtWarp& g) { return g.pos < MT_N; }

// One warp per stream.  state layout: [stream][2][625] (numpy, python).
__global__ void __launch_bounds__(32) select_kernel(uint32_t* __restrict__ states, const int32_t* __restrict__ ncand,
                                                    const int32_t* __restrict__ cand, const int32_t* __restrict__ dom_size,
                                                    const int32_t* __restrict__ dom_off, int n_per_stream,
                                                    int64_t* __restrict__ partner, int32_t* __restrict__ ratio_k,
                                                    int ratio_lo, int ratio_width) {
  __shared__ uint32_t s_cur[MT_N], s_a[MT_N], s_b[MT_N];
  __shared__ int s_flag;
  const int stream = blockIdx.x, lane = threadIdx.x;
  for (int which = 0; which < 2; ++which) {          // 0: numpy stream -> partners; 1: python stream -> ratios
    uint32_t* st = states + ((size_t)stream * 2 + which) * RNG_STATE_WORDS;
    MtWarp g{s_cur, s_a, s_b, 0};
    for (int k = lane; k < MT_N; k += 32) {
      const uint32_t raw = st[k];
      g.a[k] = raw;
      uint32_t y = raw;                              // outputs of the current (already twisted) block, from its raw words
      y ^= y >> 11;
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= y >> 18;
      g.cur[k] = y;
    }
    g.pos = (int)st[MT_N];
    __syncwarp();
    // consumer state (lane 0): sample index, stage (numpy: 0 = domain draw, 1 = image draw)
    int i = 0, stage = 0, dom = 0;
    const int nc = ncand[stream];
    while (true) {
      if (lane == 0) {
        while (i < n_per_stream) {
          if (which == 0) {
            const int n = (stage == 0) ? nc : dom_size[dom];
            int v = 0;
            if (n > 1) {                             // n == 1: numpy returns low without consuming anything
              if (!mt_have(g)) break;
              uint32_t mask = (uint32_t)(n - 1);
              mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
              const uint32_t r = g.cur[g.pos++] & mask;
              if (r > (uint32_t)(n - 1)) continue;   // rejected: draw again
              v = (int)r;
            }
            if (stage == 0) {
              dom = cand[stream * RNG_MAX_CAND + v];
              stage = 1;
            } else {
              partner[(size_t)stream * n_per_stream + i] = (int64_t)dom_off[dom] + v;
              stage = 0;
              ++i;
            }
          } else {
            int v = 0;
            if (ratio_width > 1) {                   // CPython's _randbelow(1) also draws (getrandbits(1)); handled below
              if (!mt_have(g)) break;
              const int kbits = 32 - __clz(ratio_width);
              const uint32_t r = g.cur[g.pos++] >> (32 - kbits);
              if (r >= (uint32_t)ratio_width) continue;
              v = (int)r;
            } else {
              if (!mt_have(g)) break;
              const uint32_t r = g.cur[g.pos++] >> 31;   // bit_length(1) == 1: getrandbits(1) until < 1
              if (r >= 1u) continue;
            }
            ratio_k[(size_t)stream * n_per_stream + i] = ratio_lo + v;
            ++i;
          }
        }
        s_flag = (i < n_per_stream) ? 1 : 0;         // 1: ran out of outputs, twist and continue
      }
      __syncwarp();
      const int more = s_flag;
      __syncwarp();
      if (!more) break;
      mt_twist(g, lane);
    }
    // write the stream back: raw words of the current block + position
    const int pos = __shfl_sync(0xffffffffu, g.pos, 0);
    for (int k = lane; k < MT_N; k += 32) st[k] = g.a[k];
    if (lane == 0) st[MT_N] = (uint32_t)pos;
    __syncwarp();
  }
}

// ---- fundus mask encoding (fundus.py:227-239 + transform.py:10-14) ---------------------------------------------------
// grey > 200 -> background [0, 0]; 51..200 -> disc rim [0, 1]; <= 50 -> cup [1, 1]:  out[0] = (m <= 50), out[1] = (m <= 200)
__global__ void mask_encode_kernel(const uint8_t* __restrict__ m, float* __restrict__ out, size_t hw, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t img = i / hw, px = i - img * hw;
    const unsigned v = m[i];
    out[(img * 2) * hw + px] = v <= 50u ? 1.f : 0.f;
    out[(img * 2 + 1) * hw + px] = v <= 200u ? 1.f : 0.f;
  }
}

}  // namespace ramdsir

using namespace ramdsir;

struct ramdsir_rng {
  int n_streams, n_domains;
  uint32_t* states;      // device [n_streams][2][625]
  int32_t *ncand, *cand, *dom_size, *dom_off;
};

extern "C" {

int ramdsir_rng_state_words(void) { return RNG_STATE_WORDS; }

int ramdsir_rng_seed_host(uint32_t worker_seed, uint32_t* numpy_state, uint32_t* python_state) {
  if (!numpy_state || !python_state) {
    set_error("rng_seed_host: null pointer");
    return RAMDSIR_ERR_INVALID;
  }
  mt_init_genrand(numpy_state, worker_seed);
  mt_init_by_array(python_state, &worker_seed, 1);
  return RAMDSIR_OK;
}

int ramdsir_rng_create(ramdsir_rng_t** h, int n_streams, const uint32_t* worker_seeds, int n_domains,
                       const int32_t* domain_sizes, const int32_t* n_candidates, const int32_t* candidates) {
  if (!h || n_streams < 1 || n_domains < 1 || !worker_seeds || !domain_sizes || !n_candidates || !candidates) {
    set_error("rng_create: invalid arguments");
    return RAMDSIR_ERR_INVALID;
  }
  std::vector<int32_t> off(n_domains);
  int64_t acc = 0;
  for (int d = 0; d < n_domains; ++d) {
    if (domain_sizes[d] < 1) {
      set_error("rng_create: domain %d has %d images", d, domain_sizes[d]);
      return RAMDSIR_ERR_INVALID;
    }
    off[d] = (int32_t)acc;
    acc += domain_sizes[d];
  }
  for (int s = 0; s < n_streams; ++s) {
    if (n_candidates[s] < 1 || n_candidates[s] > RNG_MAX_CAND) {
      set_error("rng_create: stream %d has %d candidate domains (1..%d)", s, n_candidates[s], RNG_MAX_CAND);
      return RAMDSIR_ERR_INVALID;
    }
    for (int k = 0; k < n_candidates[s]; ++k)
      if (candidates[s * RNG_MAX_CAND + k] < 0 || candidates[s * RNG_MAX_CAND + k] >= n_domains) {
        set_error("rng_create: stream %d candidate %d is not a domain index", s, k);
        return RAMDSIR_ERR_INVALID;
      }
  }
  std::vector<uint32_t> st((size_t)n_streams * 2 * RNG_STATE_WORDS);
  for (int s = 0; s < n_streams; ++s)
    ramdsir_rng_seed_host(worker_seeds[s], &st[((size_t)s * 2) * RNG_STATE_WORDS], &st[((size_t)s * 2 + 1) * RNG_STATE_WORDS]);
  ramdsir_rng* r = new ramdsir_rng();
  memset(r, 0, sizeof(*r));
  r->n_streams = n_streams;
  r->n_domains = n_domains;
  cudaError_t e = cudaMalloc(&r->states, st.size() * 4);
  if (e == cudaSuccess) e = cudaMalloc(&r->ncand, sizeof(int32_t) * n_streams);
  if (e == cudaSuccess) e = cudaMalloc(&r->cand, sizeof(int32_t) * n_streams * RNG_MAX_CAND);
  if (e == cudaSuccess) e = cudaMalloc(&r->dom_size, sizeof(int32_t) * n_domains);
  if (e == cudaSuccess) e = cudaMalloc(&r->dom_off, sizeof(int32_t) * n_domains);
  if (e == cudaSuccess) e = cudaMemcpy(r->states, st.data(), st.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(r->ncand, n_candidates, sizeof(int32_t) * n_streams, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(r->cand, candidates, sizeof(int32_t) * n_streams * RNG_MAX_CAND, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(r->dom_size, domain_sizes, sizeof(int32_t) * n_domains, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(r->dom_off, off.data(), sizeof(int32_t) * n_domains, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();     // creation is a setup call; the pageable uploads are complete on return
  if (e != cudaSuccess) {
    ramdsir_rng_destroy(r);
    return cuda_fail(e, "rng_create");
  }
  *h = r;
  return RAMDSIR_OK;
}

void ramdsir_rng_destroy(ramdsir_rng_t* r) {
  if (!r) return;
  cudaFree(r->states); cudaFree(r->ncand); cudaFree(r->cand); cudaFree(r->dom_size); cudaFree(r->dom_off);
  delete r;
}

int ramdsir_rng_draw(ramdsir_rng_t* r, int n_per_stream, int64_t* partner_idx_dev, int32_t* ratio_k_dev, int ratio_lo,
                     int ratio_hi, void* stream) {
  if (!r || n_per_stream < 0 || !partner_idx_dev || !ratio_k_dev || ratio_hi < ratio_lo) {
    set_error("rng_draw: invalid arguments");
    return RAMDSIR_ERR_INVALID;
  }
  if (n_per_stream == 0) return RAMDSIR_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  fused_trail_reset(st);
  select_kernel<<<r->n_streams, 32, 0, st>>>(r->states, r->ncand, r->cand, r->dom_size, r->dom_off, n_per_stream,
                                              partner_idx_dev, ratio_k_dev, ratio_lo, ratio_hi - ratio_lo + 1);
  count_launch();
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

int ramdsir_rng_get_state(ramdsir_rng_t* r, int stream_index, uint32_t* numpy_state, uint32_t* python_state) {
  if (!r || stream_index < 0 || stream_index >= r->n_streams || !numpy_state || !python_state) {
    set_error("rng_get_state: invalid arguments");
    return RAMDSIR_ERR_INVALID;
  }
  RD_CUDA(cudaDeviceSynchronize());
  RD_CUDA(cudaMemcpy(numpy_state, r->states + ((size_t)stream_index * 2) * RNG_STATE_WORDS, RNG_STATE_WORDS * 4, cudaMemcpyDeviceToHost));
  RD_CUDA(cudaMemcpy(python_state, r->states + ((size_t)stream_index * 2 + 1) * RNG_STATE_WORDS, RNG_STATE_WORDS * 4, cudaMemcpyDeviceToHost));
  return RAMDSIR_OK;
}

int ramdsir_encode_fundus_mask_u8(const uint8_t* mask_dev, float* out_dev, int B, int H, int W, void* stream) {
  if (B < 0 || H < 1 || W < 1 || (B > 0 && (!mask_dev || !out_dev))) {
    set_error("encode_fundus_mask: invalid arguments");
    return RAMDSIR_ERR_INVALID;
  }
  if (B == 0) return RAMDSIR_OK;
  const size_t hw = (size_t)H * W, total = hw * B;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  fused_trail_reset(st);
  const int threads = 256;
  size_t blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mask_encode_kernel<<<(unsigned)blocks, threads, 0, st>>>(mask_dev, out_dev, hw, total);
  count_launch();
  RD_CUDA(cudaGetLastError());
  return RAMDSIR_OK;
}

}  // extern "C"
