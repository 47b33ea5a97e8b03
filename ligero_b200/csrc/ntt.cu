// Reed-Solomon row encoding over BN254 Fr for sm_100a  (replaces src/ligero/mod.rs:521-533, 998-1008).
//
// Reference schedule: per row iFFT_k, then zero-pad to n = rho_inv*k and FFT_n.  Here (exactly the same
// values): U[i][rho_inv*c + s] = p_i(g^s w_k^c) is the size-k NTT of the coefficients
// scaled by g^(s*idx), g = w_n; coset 0 is the message itself.  Per row:
//     DIF inverse NTT (natural in -> bit-reversed coefficients, no reordering pass)
//     x (rho_inv-1):  scale by table[s][pos] = g^(s*bitrev(pos))/k   then DIT NTT (bit-reversed in -> natural)
// A CTA keeps 2^LOG_E elements in shared memory (two 16-byte half planes, XOR-swizzled so every access
// pattern of the radix-8 passes is bank-conflict free); each thread holds 8 elements in registers and
// does three butterfly stages per shared-memory round trip.  Rows longer than 2^LOG_E get their top
// stages from register-only radix-2/4/8 passes over global memory (fully coalesced: consecutive threads
// touch consecutive 32-byte elements).  All-zero rows (about half of the witness matrix) are
// detected at load time and short-circuited to zero stores.
#include <cstdlib>

#include "fr_host.h"
#include "fr_lazy.cuh"
#include "lg_internal.h"

namespace lg {

// ------------------------------------------------------------------------------------------------
// element load/store helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fr fr_pack(const uint4& a, const uint4& b) {
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ Fr ldg_fr(const Fr* p) {  // read-only path (tables)
  const uint4* q = reinterpret_cast<const uint4*>(p);
  return fr_pack(__ldg(q), __ldg(q + 1));
}
__device__ __forceinline__ FrTw ldg_tw(const FrTw* p) {  // table entry: constant + quotient multiplier
  const uint4* q = reinterpret_cast<const uint4*>(p);
  FrTw t;
  t.w = fr_pack(__ldg(q), __ldg(q + 1));
  t.p = fr_pack(__ldg(q + 2), __ldg(q + 3));
  return t;
}
__device__ __forceinline__ Fr ld_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  return fr_pack(q[0], q[1]);
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& x) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ uint32_t fr_or(const Fr& x) {
  return x.v[0] | x.v[1] | x.v[2] | x.v[3] | x.v[4] | x.v[5] | x.v[6] | x.v[7];
}

// shared memory: element i lives at 16-byte slot swz(i) of the lo plane and of the hi plane
__device__ __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 3) & 7u); }
__device__ __forceinline__ Fr lds_fr(const uint4* lo, const uint4* hi, uint32_t i) {
  const uint32_t p = swz(i);
  return fr_pack(lo[p], hi[p]);
}
__device__ __forceinline__ void sts_fr(uint4* lo, uint4* hi, uint32_t i, const Fr& x) {
  const uint32_t p = swz(i);
  lo[p] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  hi[p] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

// ------------------------------------------------------------------------------------------------
// radix-2^R butterfly network on the R stages s .. s+R-1 of a size-2^q transform.
// x[e] is the element whose index has bits (s..s+R-1) = e and low s bits = t_lo.
//   DIT (forward): v = x1 * w^j ; (x0, x1) = (x0 + v, x0 - v)          stages ascending
//   DIF (inverse): (x0, x1) = (x0 + x1, (x0 - x1) * w^-j)              stages descending
// with j = index mod 2^stage and twiddle W[j << (q - stage - 1)], W the order-2^q table.
// Values stay lazily reduced between stages (fr_lazy.cuh): < 4r + d forward, < 2r + d inverse; whoever
// stores a FINISHED element applies fr_normalize, so what reaches U is the canonical representative.
// ------------------------------------------------------------------------------------------------
// LATE: load each half of a table entry only when the product needs it (fr_mul_shoup_ld) instead of holding the
// whole 16-register entry across the butterflies that share it -- for the register-starved shared-memory passes
template <int R, bool DIF, bool LATE = false>
__device__ __forceinline__ void butterflies(Fr (&x)[1 << R], uint32_t t_lo, int s, int q,
                                            const FrTw* __restrict__ W) {
#pragma unroll
  for (int bb = 0; bb < R; bb++) {
    const int b = DIF ? (R - 1 - bb) : bb;
    const int shift = q - (s + b) - 1;
#pragma unroll
    for (int el = 0; el < (1 << b); el++) {
      const uint32_t j = t_lo | ((uint32_t)el << s);
      if (j == 0) {  // twiddle 1: no multiplication (all of stage 0, half of stage 1, ...)
#pragma unroll
        for (int eh = 0; eh < (1 << (R - 1 - b)); eh++) {
          const int e0 = el | (eh << (b + 1)), e1 = e0 | (1 << b);
          if (!DIF) lz_bfly_dit1(x[e0], x[e1]);
          else lz_bfly_dif1(x[e0], x[e1]);
        }
      } else if (LATE) {
        const FrTw* wp = W + ((size_t)j << shift);
#pragma unroll
        for (int eh = 0; eh < (1 << (R - 1 - b)); eh++) {
          const int e0 = el | (eh << (b + 1)), e1 = e0 | (1 << b);
          if (!DIF) lz_bfly_dit_ld(x[e0], x[e1], wp);
          else lz_bfly_dif_ld(x[e0], x[e1], wp);
        }
      } else {
        const FrTw w = ldg_tw(W + ((size_t)j << shift));
#pragma unroll
        for (int eh = 0; eh < (1 << (R - 1 - b)); eh++) {
          const int e0 = el | (eh << (b + 1)), e1 = e0 | (1 << b);
          if (!DIF) lz_bfly_dit(x[e0], x[e1], w);
          else lz_bfly_dif(x[e0], x[e1], w);
        }
      }
    }
  }
}

// one radix-2^R pass over E elements held in shared memory (src planes -> dst planes; may alias)
template <int R, bool DIF, bool SCALE>
__device__ __forceinline__ void smem_pass(const uint4* slo, const uint4* shi, uint4* dlo, uint4* dhi, int s, int q,
                                          const FrTw* __restrict__ W, const FrTw* __restrict__ scale, uint32_t col_base,
                                          uint32_t col_mask, int E, int NT) {
#pragma unroll 2
  for (int g = threadIdx.x; g < (E >> R); g += NT) {
    const uint32_t t_lo = g & ((1u << s) - 1u), t_hi = (uint32_t)g >> s;
    const uint32_t base = (t_hi << (s + R)) | t_lo;
    Fr x[1 << R];
#pragma unroll
    for (int e = 0; e < (1 << R); e++) {
      const uint32_t idx = base | ((uint32_t)e << s);
      x[e] = lds_fr(slo, shi, idx);
      if (SCALE) x[e] = fr_mul_shoup_ld(x[e], scale + ((col_base + idx) & col_mask));
    }
    butterflies<R, DIF, true>(x, t_lo, s, q, W);
#pragma unroll
    for (int e = 0; e < (1 << R); e++) sts_fr(dlo, dhi, base | ((uint32_t)e << s), x[e]);
  }
}

template <int MAXR, bool DIF, bool SCALE>
__device__ __forceinline__ void smem_pass_r(int r, const uint4* slo, const uint4* shi, uint4* dlo, uint4* dhi, int s,
                                            int q, const FrTw* W, const FrTw* scale, uint32_t col_base, uint32_t col_mask,
                                            int E, int NT) {
  if (MAXR >= 3 && r == 3) smem_pass<(MAXR >= 3 ? 3 : 2), DIF, SCALE>(slo, shi, dlo, dhi, s, q, W, scale, col_base, col_mask, E, NT);
  else if (r == 2) smem_pass<2, DIF, SCALE>(slo, shi, dlo, dhi, s, q, W, scale, col_base, col_mask, E, NT);
  else smem_pass<1, DIF, SCALE>(slo, shi, dlo, dhi, s, q, W, scale, col_base, col_mask, E, NT);
}

// stage grouping: radix 8 wherever possible, never a lone radix-2 after radix-8 when 4 stages remain
__host__ __device__ __forceinline__ int pass_radix(int remaining, int maxr = 3) {
  if (maxr == 2) return remaining < 2 ? remaining : 2;
  return remaining == 4 ? 2 : (remaining < 3 ? remaining : 3);
}

struct LocalArgs {
  const Fr* in;         // rows*k elements (message, or output of the strided DIF passes)
  Fr* out;              // MODE 0: U planes base;  MODE 1: coefficient output (natural order)
  unsigned long long total;         // rows * k
  unsigned long long plane_stride;  // rows * k
  int q;                // log2 k
  int l;                // local stages = min(q, LOG_E)
  int rho;              // cosets (MODE 0)
  Fr* plane0;           // MODE 0: if non-null, also store the input here (only when `in` is the message)
  const FrTw* w_fwd;    // order-2^l tables (compact: every local stage indexes a dense 2^(l-1)-entry array)
  const FrTw* w_inv;
  const FrTw* scale;
  const uint4* scale4;  // ntt_persist_kernel: SoA copy of `scale` (NttTables::scale4)
  FrTw kinv;            // MODE 1: 1/k
  int final;            // MODE 0: the coset values leaving this kernel are finished codeword elements (q == l)
  int mapped;           // MODE 0: 1 = finished elements (and the plane-0 copy) go through `map`
  int copy0;            // MODE 0, mapped: also store the input as plane 0
  int plain0;           // MODE 0: the plane-0 copy is stored as plain integers (from_mont), like the coset planes
  OutMap map;
};

// mapped store of element `f` (flat index into the local rows x k array) of coset plane s
__device__ __forceinline__ void st_mapped(const LocalArgs& a, uint32_t s, unsigned long long f, const Fr& x) {
  const uint32_t row = (uint32_t)(f >> a.q), col = (uint32_t)f & ((1u << a.q) - 1u);
  st_fr(outmap_ptr(a.map, s, outmap_row(a.map, row), col), x);
}

// MODE 0: encode (iNTT tail + all cosets).  MODE 1: iNTT tail only, scaled, natural-order output.
template <int LOG_E, int MAXR, int MINB, int MODE, int NT = (1 << (LOG_E - MAXR))>
__global__ void __launch_bounds__(NT, MINB) ntt_local_kernel(const __grid_constant__ LocalArgs a) {
  constexpr int E = 1 << LOG_E;
  extern __shared__ uint4 smem[];
  uint4 *Alo = smem, *Ahi = smem + E, *Blo = smem + 2 * E, *Bhi = smem + 3 * E;
  const unsigned long long f0 = (unsigned long long)blockIdx.x * E;
  const uint32_t col_mask = (1u << a.q) - 1u;
  const uint32_t col_base = (uint32_t)(f0 & col_mask);

  uint32_t nz = 0;
  for (int i = threadIdx.x; i < E; i += NT) {
    const unsigned long long f = f0 + i;
    Fr x = fr_zero();
    if (f < a.total) x = ld_fr(a.in + f);
    nz |= fr_or(x);
    sts_fr(Alo, Ahi, i, x);
    if (MODE == 0 && f < a.total) {
      if (a.mapped ? a.copy0 != 0 : a.plane0 != nullptr) {
        const Fr x0 = a.plain0 ? fr_from_mont(x) : x;
        if (a.mapped) st_mapped(a, 0, f, x0);
        else st_fr(a.plane0 + f, x0);
      }
    }
  }
  const int any = __syncthreads_or(nz != 0);
  if (!any) {  // all-zero rows: codeword / coefficients are zero
    const Fr z = fr_zero();
    if (MODE == 0) {
      for (int s = 1; s < a.rho; s++)
        for (int i = threadIdx.x; i < E; i += NT)
          if (f0 + i < a.total) {
            if (a.mapped) st_mapped(a, s, f0 + i, z);
            else st_fr(a.out + (s - 1) * a.plane_stride + f0 + i, z);
          }
    } else {
      for (int i = threadIdx.x; i < E; i += NT)
        if (f0 + i < a.total) st_fr(a.out + f0 + i, z);
    }
    return;
  }

  // inverse transform tail: DIF stages l-1 .. 0 in place in A
  for (int top = a.l; top > 0;) {
    const int r = pass_radix(top, MAXR), s = top - r;
    smem_pass_r<MAXR, true, false>(r, Alo, Ahi, Alo, Ahi, s, a.l, a.w_inv, nullptr, 0, 0, E, NT);
    __syncthreads();
    top = s;
  }

  if (MODE == 1) {
    for (int i = threadIdx.x; i < E; i += NT) {
      const unsigned long long f = f0 + i;
      if (f < a.total) {
        const uint32_t col = (uint32_t)(f & col_mask);
        const uint32_t nat = __brev(col) >> (32 - a.q);
        st_fr(a.out + (f - col) + nat, fr_normalize_2r(fr_mul_shoup(lds_fr(Alo, Ahi, i), a.kinv)));
      }
    }
    return;
  }

  // cosets 1 .. rho-1: scale (first pass, A -> B) then DIT stages 0 .. l-1 in B
  for (int cs = 1; cs < a.rho; cs++) {
    const FrTw* sc = a.scale + (size_t)(cs - 1) * ((size_t)1 << a.q);
    for (int s = 0; s < a.l;) {
      const int r = pass_radix(a.l - s, MAXR);
      if (s == 0) smem_pass_r<MAXR, false, true>(r, Alo, Ahi, Blo, Bhi, 0, a.l, a.w_fwd, sc, col_base, col_mask, E, NT);
      else smem_pass_r<MAXR, false, false>(r, Blo, Bhi, Blo, Bhi, s, a.l, a.w_fwd, nullptr, 0, 0, E, NT);
      __syncthreads();
      s += r;
    }
    Fr* dst = a.out + (cs - 1) * a.plane_stride + f0;
    for (int i = threadIdx.x; i < E; i += NT)
      if (f0 + i < a.total) {
        Fr v = lds_fr(Blo, Bhi, i);
        if (a.final) v = fr_normalize(v);
        if (a.mapped) st_mapped(a, cs, f0 + i, v);
        else st_fr(dst + i, v);
      }
    __syncthreads();
  }
}

// register-only radix-2^R pass over global memory for the stages s..s+R-1 (s >= local size) of every row
template <int R, bool DIF, int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB) ntt_global_pass_kernel(const Fr* in, Fr* out, Fr* copy_out, int q, int s,
                                                              const FrTw* __restrict__ W, int final, int copy_plain) {
  const uint32_t g = blockIdx.y * blockDim.x + threadIdx.x;
  if (g >= (1u << (q - R))) return;
  const size_t off = (size_t)blockIdx.x << q;
  const uint32_t t_lo = g & ((1u << s) - 1u), t_hi = g >> s;
  const uint32_t base = (t_hi << (s + R)) | t_lo;
  Fr x[1 << R];
  uint32_t nz = 0;
#pragma unroll
  for (int e = 0; e < (1 << R); e++) {
    x[e] = ld_fr(in + off + (base | ((uint32_t)e << s)));
    nz |= fr_or(x[e]);
  }
  if (copy_out) {
#pragma unroll
    for (int e = 0; e < (1 << R); e++)
      st_fr(copy_out + off + (base | ((uint32_t)e << s)), (copy_plain && nz) ? fr_from_mont(x[e]) : x[e]);
  }
  if (nz == 0 && in == out) return;  // zeros stay zeros
  if (nz != 0) {
    butterflies<R, DIF>(x, t_lo, s, q, W);
    if (final) {
#pragma unroll
      for (int e = 0; e < (1 << R); e++) x[e] = fr_normalize(x[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < (1 << R); e++) st_fr(out + off + (base | ((uint32_t)e << s)), x[e]);
}

// the same pass with mapped stores (multi-GPU): COPY0 -> the loaded input is plane 0 of row blockIdx.x and
// the transformed values go to the local `out`; otherwise `in` holds rows_per_plane rows per coset plane
// (plane 1 first) and the transformed values are final and go through the map
template <int R, bool DIF, bool COPY0>
__global__ void __launch_bounds__(128, 3) ntt_global_pass_mapped_kernel(const Fr* in, Fr* out, int q, int s,
                                                                     const FrTw* __restrict__ W, const __grid_constant__ OutMap map,
                                                                     uint32_t rows_per_plane, int copy_plain) {
  const uint32_t g = blockIdx.y * blockDim.x + threadIdx.x;
  if (g >= (1u << (q - R))) return;
  const size_t off = (size_t)blockIdx.x << q;
  const uint32_t t_lo = g & ((1u << s) - 1u), t_hi = g >> s;
  const uint32_t base = (t_hi << (s + R)) | t_lo;
  const uint32_t plane = COPY0 ? 0u : 1u + blockIdx.x / rows_per_plane;
  const uint32_t grow = outmap_row(map, COPY0 ? blockIdx.x : blockIdx.x % rows_per_plane);
  Fr x[1 << R];
  uint32_t nz = 0;
#pragma unroll
  for (int e = 0; e < (1 << R); e++) {
    x[e] = ld_fr(in + off + (base | ((uint32_t)e << s)));
    nz |= fr_or(x[e]);
  }
  if (COPY0) {
#pragma unroll
    for (int e = 0; e < (1 << R); e++)
      st_fr(outmap_ptr(map, 0, grow, base | ((uint32_t)e << s)), (copy_plain && nz) ? fr_from_mont(x[e]) : x[e]);
  }
  if (nz != 0) {
    butterflies<R, DIF>(x, t_lo, s, q, W);
    if (!COPY0) {  // finished codeword elements
#pragma unroll
      for (int e = 0; e < (1 << R); e++) x[e] = fr_normalize(x[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < (1 << R); e++) {
    if (COPY0) st_fr(out + off + (base | ((uint32_t)e << s)), x[e]);
    else st_fr(outmap_ptr(map, plane, grow, base | ((uint32_t)e << s)), x[e]);
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent shared-memory encoder for rows of >= 1024 elements (the MODE 0 work of ntt_local_kernel).
//
// One CTA per SM, GROUPS independent groups of 256 threads (8 warps), each walking 1024-element chunks with its own
// named barrier and its own 64 KiB of shared memory (coefficients A, working copy B, two 16-byte half planes each).
// What the one-CTA-per-chunk kernel paid per product and this one does not:
//   * twiddles come from ONE shared-memory copy of the order-1024 table per SM (513 entries as four 16-byte planes,
//     XOR-swizzled: every access pattern of the five radix-4 passes is conflict free) instead of four scattered
//     64-byte global loads per product through a 32 KiB L1 -- the late passes were bound by the LSU, not the multiplier;
//     the inverse transform reads the same table backwards: omega^-i = -omega^(512-i), (X-Y) omega^-i = (Y-X) omega^(512-i);
//   * the data swizzle is conflict free for radix-4 (the radix-8 swizzle of ntt_local_kernel is 2-way conflicted in
//     the pass over stages 2,3);
//   * the first inverse pass takes its four elements straight from global memory and the last forward pass stores its
//     four finished elements straight to global memory (both coalesced: the elements of thread t are t + 256 e), saving
//     one shared-memory round trip and one barrier per transform;
//   * the coset scale factors are read from the SoA copy (NttTables::scale4), consecutive lanes -> consecutive slots;
//   * unit twiddles are skipped only where the whole warp skips (stages 0 and half of stage 1): per-lane skipping in the
//     later passes made every warp execute both sides of the branch.
// ------------------------------------------------------------------------------------------------
constexpr int kLogPE = 10;
constexpr int kPE = 1 << kLogPE;   // elements per chunk
constexpr int kPT = kPE / 4;       // threads per group
constexpr int kTwSlots = 520;      // 513 twiddles, padded
constexpr size_t kPersistTwBytes = 4 * kTwSlots * sizeof(uint4);
constexpr size_t kPersistGroupBytes = 4 * kPE * sizeof(uint4);

__device__ __forceinline__ uint32_t swz4(uint32_t i) { return i ^ ((i >> 3) & 1u) ^ (((i >> 4) & 1u) * 6u); }
__device__ __forceinline__ uint32_t tsw(uint32_t e) { return e ^ ((e >> 3) & 7u) ^ ((e >> 6) & 7u); }
__device__ __forceinline__ Fr lds4(const uint4* lo, uint32_t i) {
  const uint32_t p = swz4(i);
  return fr_pack(lo[p], lo[kPE + p]);
}
__device__ __forceinline__ void sts4(uint4* lo, uint32_t i, const Fr& x) {
  const uint32_t p = swz4(i);
  lo[p] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  lo[kPE + p] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ void group_sync(uint32_t id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ bool group_any(uint32_t id, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, 256, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"((uint32_t)pred), "r"(id)
      : "memory");
  return r != 0;
}
// Y * omega^e with the twiddle planes at shared address tw (entry e at slot tsw(e))
__device__ __forceinline__ Fr mul_tw_sm(const Fr& y, uint32_t tw, uint32_t e) {
  const uint32_t a = tw + (tsw(e) << 4);
  return fr_mul_shoup_sm(y, a, a + kTwSlots * 16, a + 2 * kTwSlots * 16, a + 3 * kTwSlots * 16);
}
__device__ __forceinline__ void bfly_dit_sm(Fr& X, Fr& Y, uint32_t tw, uint32_t e) {
  lz_csub2r(X);
  const Fr T = mul_tw_sm(Y, tw, e);
  Y = lz_add(X, lz_2r_minus(T));
  X = lz_add(X, T);
}
// inverse butterfly with omega^-i read as -omega^(512-i): (X, Y) -> (X + Y, (Y - X) omega^(512-i))
__device__ __forceinline__ void bfly_dif_sm(Fr& X, Fr& Y, uint32_t tw, uint32_t i) {
  const Fr D = lz_add(Y, lz_3r_minus(X));
  X = lz_add(X, Y);
  lz_csub2r(X);
  Y = mul_tw_sm(D, tw, 512u - i);
}
// the two stages s, s+1 of a size-1024 transform on the four elements of one thread (x[e]: index bits s, s+1 = e)
template <bool DIF>
__device__ __forceinline__ void bfly4_sm(Fr (&x)[4], uint32_t t_lo, int s, uint32_t tw) {
  if (!DIF) {
    if (s == 0) {
      lz_bfly_dit1(x[0], x[1]);
      lz_bfly_dit1(x[2], x[3]);
      lz_bfly_dit1(x[0], x[2]);
      bfly_dit_sm(x[1], x[3], tw, 256u);
    } else {
      const uint32_t e0 = t_lo << (9 - s);
      bfly_dit_sm(x[0], x[1], tw, e0);
      bfly_dit_sm(x[2], x[3], tw, e0);
      const uint32_t e1 = t_lo << (8 - s);
      bfly_dit_sm(x[0], x[2], tw, e1);
      bfly_dit_sm(x[1], x[3], tw, e1 | 256u);
    }
  } else {
    if (s == 0) {
      lz_bfly_dif1(x[0], x[2]);
      bfly_dif_sm(x[1], x[3], tw, 256u);
      lz_bfly_dif1(x[0], x[1]);
      lz_bfly_dif1(x[2], x[3]);
    } else {
      const uint32_t e1 = t_lo << (8 - s);
      bfly_dif_sm(x[0], x[2], tw, e1);
      bfly_dif_sm(x[1], x[3], tw, e1 | 256u);
      const uint32_t e0 = t_lo << (9 - s);
      bfly_dif_sm(x[0], x[1], tw, e0);
      bfly_dif_sm(x[2], x[3], tw, e0);
    }
  }
}
template <bool DIF>
__device__ __forceinline__ void pass4_inplace(uint4* buf, int s, uint32_t t, uint32_t tw) {
  const uint32_t t_lo = t & ((1u << s) - 1u), base = ((t >> s) << (s + 2)) | t_lo;
  Fr x[4];
#pragma unroll
  for (int e = 0; e < 4; e++) x[e] = lds4(buf, base | ((uint32_t)e << s));
  bfly4_sm<DIF>(x, t_lo, s, tw);
#pragma unroll
  for (int e = 0; e < 4; e++) sts4(buf, base | ((uint32_t)e << s), x[e]);
}

// CAPPED: at most 80 registers per thread (two groups: 40 K of the SM's 64 K registers; four 64-thread CTAs of the column-hash
// kernel at 80 registers take 20 K more -- with 88 registers the sum was exactly 64 K and the hash CTAs never became co-resident;
// the multi-GPU block pipeline of capi_shard.cu)
template <int GROUPS, bool CAPPED>
__global__ void __launch_bounds__(GROUPS* kPT, 1) __maxnreg__(CAPPED ? 80 : 255)
    ntt_persist_kernel(const __grid_constant__ LocalArgs a) {
  extern __shared__ uint4 smem[];
  // one copy of the order-1024 twiddles per SM
  for (uint32_t e = threadIdx.x; e <= 512; e += GROUPS * kPT) {
    const uint4* src = reinterpret_cast<const uint4*>(a.w_fwd + e);
    const uint32_t p = tsw(e);
#pragma unroll
    for (int pl = 0; pl < 4; pl++) smem[pl * kTwSlots + p] = __ldg(src + pl);
  }
  __syncthreads();
  const uint32_t t = threadIdx.x & (kPT - 1), gi = threadIdx.x / kPT, bar = 1 + gi;
  const uint32_t tw = (uint32_t)__cvta_generic_to_shared(smem);
  uint4* A = smem + 4 * kTwSlots + (size_t)gi * 4 * kPE;
  uint4* B = A + 2 * kPE;
  const unsigned long long nchunks = a.total >> kLogPE;
  const uint32_t col_mask = (1u << a.q) - 1u;
  const bool copy0 = a.mapped ? a.copy0 != 0 : a.plane0 != nullptr;
  for (unsigned long long ch = (unsigned long long)blockIdx.x * GROUPS + gi; ch < nchunks;
       ch += (unsigned long long)gridDim.x * GROUPS) {
    const unsigned long long f0 = ch << kLogPE;
    Fr x[4];
    uint32_t nz = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
      x[e] = ld_fr(a.in + f0 + t + kPT * e);
      nz |= fr_or(x[e]);
    }
    if (copy0) {
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const Fr x0 = a.plain0 ? fr_from_mont(x[e]) : x[e];
        if (a.mapped) st_mapped(a, 0, f0 + t + kPT * e, x0);
        else st_fr(a.plane0 + f0 + t + kPT * e, x0);
      }
    }
    // (the barrier also orders the previous chunk's last reads of B before this chunk's first writes)
    if (!group_any(bar, nz != 0)) {  // all-zero chunk: the codeword is zero
      const Fr z = fr_zero();
      for (int cs = 1; cs < a.rho; cs++)
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const unsigned long long f = f0 + t + kPT * e;
          if (a.mapped) st_mapped(a, cs, f, z);
          else st_fr(a.out + (cs - 1) * a.plane_stride + f, z);
        }
      continue;
    }
    // inverse tail: DIF stages 9..0; the first pass works on the registers just loaded (thread t holds t + 256 e)
    bfly4_sm<true>(x, t, 8, tw);
#pragma unroll
    for (int e = 0; e < 4; e++) sts4(A, t | ((uint32_t)e << 8), x[e]);
    group_sync(bar);
#pragma unroll 1
    for (int s = 6; s >= 0; s -= 2) {
      pass4_inplace<true>(A, s, t, tw);
      group_sync(bar);
    }
    // cosets 1 .. rho-1: scale while loading A, DIT stages 0..9 in B, the last pass stores to global memory
    const uint4* sc = a.scale4 + (size_t)((uint32_t)(f0 & col_mask) >> kLogPE) * (4 * kPE) + t;
    const size_t sc_stride = ((size_t)1 << (a.q - kLogPE)) * (4 * kPE);
#pragma unroll 1
    for (int cs = 1; cs < a.rho; cs++, sc += sc_stride) {
#pragma unroll
      for (int e = 0; e < 4; e++) x[e] = fr_mul_shoup_g4(lds4(A, (t << 2) | (uint32_t)e), sc + e * kPT, kPE);
      bfly4_sm<false>(x, 0, 0, tw);
#pragma unroll
      for (int e = 0; e < 4; e++) sts4(B, (t << 2) | (uint32_t)e, x[e]);
      group_sync(bar);
#pragma unroll 1
      for (int s = 2; s <= 6; s += 2) {
        pass4_inplace<false>(B, s, t, tw);
        group_sync(bar);
      }
#pragma unroll
      for (int e = 0; e < 4; e++) x[e] = lds4(B, t | ((uint32_t)e << 8));
      bfly4_sm<false>(x, t, 8, tw);
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const unsigned long long f = f0 + t + kPT * e;
        const Fr v = a.final ? fr_normalize(x[e]) : x[e];
        if (a.mapped) st_mapped(a, cs, f, v);
        else st_fr(a.out + (cs - 1) * a.plane_stride + f, v);
      }
      group_sync(bar);  // B is rewritten by the next coset's first pass
    }
  }
}

// SoA copy of the coset scale tables for ntt_persist_kernel (NttTables::scale4)
__global__ void scale_soa_kernel(const FrTw* __restrict__ scale, uint4* __restrict__ out, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // entry index: (coset-1)*k + position
  if (i >= count) return;
  const size_t chunk = i >> kLogPE;
  const uint32_t idx = (uint32_t)i & (kPE - 1), slot = (idx & 3u) * kPT + (idx >> 2);
  const uint4* src = reinterpret_cast<const uint4*>(scale + i);
#pragma unroll
  for (int pl = 0; pl < 4; pl++) out[chunk * (4 * kPE) + (size_t)pl * kPE + slot] = src[pl];
}

// table generation: p = floor(w 2^256 / r) for every entry
__global__ void fill_quotients_kernel(FrTw* t, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr p = fr_shoup_quotient(ld_fr(&t[i].w));
  st_fr(&t[i].p, p);
}

// ------------------------------------------------------------------------------------------------
// host side: tables and launchers
// ------------------------------------------------------------------------------------------------
static uint32_t bitrev_host(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}

int get_tables(Ctx* ctx, int log_k, int rho_inv, const NttTables** out, bool plain) {
  auto key = std::make_pair(log_k, rho_inv | (plain ? 1 << 16 : 0));
  auto it = ctx->tables.find(key);
  if (it != ctx->tables.end()) {
    *out = &it->second;
    return OK;
  }
  if (log_k < 1 || log_k > 26) return set_error(ctx, ERR_INVALID, "log_k out of range");
  int log_rho = 0;
  while ((1 << log_rho) < rho_inv) log_rho++;
  if ((1 << log_rho) != rho_inv || rho_inv < 1 || log_k + log_rho > 28)
    return set_error(ctx, ERR_INVALID, "rho_inv must be a power of two with n <= 2^28");
  const size_t k = (size_t)1 << log_k, half = k / 2 ? k / 2 : 1;
  // Montgomery-form powers on the host, then every entry becomes {plain integer, 0}; the quotient
  // multipliers are filled in on the device (fill_quotients_kernel)
  // w_fwd carries one extra entry, omega^(k/2) = -1: the persistent encoder reads its inverse twiddles as
  // omega^-i = -omega^(k/2 - i) from the forward table, i = 0 included
  std::vector<FrTw> tab(2 * half + 1 + (size_t)(rho_inv - 1) * k);
  FrTw* wf = tab.data();
  FrTw* wi = wf + half + 1;
  FrTw* sc = wi + half;
  const Fr w = fr_root_of_unity(log_k), winv = fr_inv(w);
  Fr a = fr_one(), b = fr_one();
  for (size_t i = 0; i < half; i++) {
    wf[i].w = fr_from_mont(a);
    wi[i].w = fr_from_mont(b);
    wf[i].p = wi[i].p = fr_zero();
    a = fr_mul(a, w);
    b = fr_mul(b, winv);
  }
  wf[half].w = fr_from_mont(a);  // omega^(k/2) (= r - 1 for k >= 2)
  wf[half].p = fr_zero();
  const Fr g = fr_root_of_unity(log_k + log_rho);
  const Fr kinv = fr_inv(fr_from_u64(k));
  std::vector<Fr> pw(k);
  Fr gs = fr_one();
  for (int s = 1; s < rho_inv; s++) {
    gs = fr_mul(gs, g);  // g^s
    pw[0] = kinv;
    for (size_t i = 1; i < k; i++) pw[i] = fr_mul(pw[i - 1], gs);
    for (size_t pos = 0; pos < k; pos++) {
      FrTw& e = sc[(size_t)(s - 1) * k + pos];
      e.w = fr_from_mont(pw[bitrev_host((uint32_t)pos, log_k)]);
      if (plain) e.w = fr_from_mont(e.w);  // extra R^-1: Montgomery-form input -> plain-integer coset values
      e.p = fr_zero();
    }
  }
  NttTables t;
  t.log_k = log_k;
  t.rho_inv = rho_inv;
  t.kinv.w = fr_from_mont(kinv);
  t.kinv.p = fr_shoup_quotient(t.kinv.w);
  FrTw* dev;
  LG_CUDA(ctx, cudaMalloc(&dev, tab.size() * sizeof(FrTw)));
  LG_CUDA(ctx, cudaMemcpyAsync(dev, tab.data(), tab.size() * sizeof(FrTw), cudaMemcpyHostToDevice, ctx->stream));
  fill_quotients_kernel<<<(unsigned)((tab.size() + 127) / 128), 128, 0, ctx->stream>>>(dev, tab.size());
  ctx->launches++;
  LG_CUDA(ctx, cudaGetLastError());
  LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host vector dies at scope exit
  t.w_fwd = dev;
  t.w_inv = dev + half + 1;
  t.scale = dev + 2 * half + 1;
  if (log_k >= kLogPE && rho_inv > 1) {
    const size_t cnt = (size_t)(rho_inv - 1) * k;
    LG_CUDA(ctx, cudaMalloc(&t.scale4, cnt * sizeof(FrTw)));
    scale_soa_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(t.scale, t.scale4, cnt);
    ctx->launches++;
    LG_CUDA(ctx, cudaGetLastError());
    LG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  auto ins = ctx->tables.emplace(key, t);
  *out = &ins.first->second;
  return OK;
}

constexpr int kLogE = 10;  // 1024 elements / CTA = 64 KiB of shared memory
constexpr int kMaxR = 2;   // radix-4 passes: 256 threads x 4 elements, <= 85 registers -> 3 CTAs = 24 warps per SM
constexpr int kMinB = 3;   // (measured best of the four variants below: 131 ms vs 138/164/158 ms at 16388 x 8192)

template <int R, bool DIF>
static int launch_global_pass(Ctx* ctx, const Fr* in, Fr* out, Fr* copy_out, size_t rows, int q, int s, const FrTw* W,
                              int final, int copy_plain) {
  const uint32_t groups = 1u << (q - R);
  static int mode = -1;  // LG_GP_BLOCK=256 selects the 256-thread CTAs (tuning hook)
  if (mode < 0) {
    const char* e = getenv("LG_GP_BLOCK");
    mode = (e && atoi(e) == 256) ? 1 : 0;
  }
  if (mode == 1) {
    const uint32_t bs = groups < 256 ? groups : 256;
    dim3 grid((unsigned)rows, (groups + bs - 1) / bs);
    ntt_global_pass_kernel<R, DIF, 256, 1><<<grid, bs, 0, ctx->stream>>>(in, out, copy_out, q, s, W, final, copy_plain);
  } else {
    // 128 threads x 3 CTAs per SM: room for the radix-8 register tile (8 elements + a 16-register table entry)
    const uint32_t bs = groups < 128 ? groups : 128;
    dim3 grid((unsigned)rows, (groups + bs - 1) / bs);
    ntt_global_pass_kernel<R, DIF, 128, 3><<<grid, bs, 0, ctx->stream>>>(in, out, copy_out, q, s, W, final, copy_plain);
  }
  ctx->launches++;
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}
template <bool DIF>
static int launch_global_pass_r(Ctx* ctx, int r, const Fr* in, Fr* out, Fr* copy_out, size_t rows, int q, int s,
                                const FrTw* W, int final = 0, int copy_plain = 0) {
  if (r == 3) return launch_global_pass<3, DIF>(ctx, in, out, copy_out, rows, q, s, W, final, copy_plain);
  if (r == 2) return launch_global_pass<2, DIF>(ctx, in, out, copy_out, rows, q, s, W, final, copy_plain);
  return launch_global_pass<1, DIF>(ctx, in, out, copy_out, rows, q, s, W, final, copy_plain);
}

template <int MAXR, int MINB, int MODE, int NT = (1 << (kLogE - MAXR))>
static int launch_local_v(Ctx* ctx, const LocalArgs& a) {
  constexpr int E = 1 << kLogE;
  const size_t smem = 4 * E * sizeof(uint4);
  // (per device, not per process: one process may drive several GPUs -- lg_mgpu_*)
  const void* fn = (const void*)ntt_local_kernel<kLogE, MAXR, MINB, MODE, NT>;
  if (!ctx->smem_configured.count(fn)) {
    LG_CUDA(ctx, cudaFuncSetAttribute(ntt_local_kernel<kLogE, MAXR, MINB, MODE, NT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->smem_configured.insert(fn);
  }
  const unsigned long long ctas = (a.total + E - 1) / E;
  ntt_local_kernel<kLogE, MAXR, MINB, MODE, NT><<<(unsigned)ctas, NT, smem, ctx->stream>>>(a);
  ctx->launches++;
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}

// LG_NTT_VARIANT selects alternative register/occupancy trade-offs of the same kernel (tuning hook)
static int ntt_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LG_NTT_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

template <int MODE>
static int launch_local(Ctx* ctx, const LocalArgs& a) {
  if (MODE == 0) {
    switch (ntt_variant()) {
      case 1: return launch_local_v<2, 2, MODE>(ctx, a);   // radix-4, 2 CTAs/SM (<= 128 regs)
      case 2: return launch_local_v<3, 3, MODE>(ctx, a);   // radix-8, 128 threads, 3 CTAs/SM (<= 168 regs)
      case 3: return launch_local_v<3, 2, MODE>(ctx, a);   // radix-8, 128 threads, 2 CTAs/SM
      case 4: return launch_local_v<2, 3, MODE, 128>(ctx, a);  // radix-4, 128 threads x 2 groups, 3 CTAs/SM (<= 168 regs)
      default: break;
    }
  }
  return launch_local_v<kMaxR, kMinB, MODE>(ctx, a);
}

// persistent encoder: one CTA per SM; Ctx::persist_groups picks the variant
template <int GROUPS, bool CAPPED>
static int launch_persist_g(Ctx* ctx, const LocalArgs& a) {
  const size_t smem = kPersistTwBytes + GROUPS * kPersistGroupBytes;
  const void* fn = (const void*)ntt_persist_kernel<GROUPS, CAPPED>;
  if (!ctx->smem_configured.count(fn)) {  // per device: one process may drive several GPUs
    LG_CUDA(ctx, cudaFuncSetAttribute(ntt_persist_kernel<GROUPS, CAPPED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // Ask for the largest shared-memory carve-out.  The driver otherwise picks the smallest configuration that holds
    // this CTA (164 KiB for two groups), which leaves ~2 KiB: not one CTA of a co-resident column-hash kernel (3.3 KiB
    // + 1 KiB reserved) fits, and the two kernels serialise although registers and warps are free.
    LG_CUDA(ctx, cudaFuncSetAttribute(ntt_persist_kernel<GROUPS, CAPPED>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      (int)cudaSharedmemCarveoutMaxShared));
    ctx->smem_configured.insert(fn);
  }
  const unsigned long long chunks = a.total >> kLogPE;
  unsigned long long ctas = (chunks + GROUPS - 1) / GROUPS;
  if (ctas > (unsigned long long)ctx->sm_count) ctas = ctx->sm_count;
  ntt_persist_kernel<GROUPS, CAPPED><<<(unsigned)ctas, GROUPS * kPT, smem, ctx->stream>>>(a);
  ctx->launches++;
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}
static int launch_persist(Ctx* ctx, const LocalArgs& a, int groups) {
  return groups == 2 ? launch_persist_g<2, true>(ctx, a) : launch_persist_g<3, false>(ctx, a);
}

template <int R, bool DIF, bool COPY0>
static int launch_global_pass_mapped(Ctx* ctx, const Fr* in, Fr* out, size_t rows, int q, int s, const FrTw* W,
                                     const OutMap& map, uint32_t rows_per_plane, int copy_plain) {
  const uint32_t groups = 1u << (q - R);
  const uint32_t bs = groups < 128 ? groups : 128;
  dim3 grid((unsigned)rows, (groups + bs - 1) / bs);
  ntt_global_pass_mapped_kernel<R, DIF, COPY0><<<grid, bs, 0, ctx->stream>>>(in, out, q, s, W, map, rows_per_plane, copy_plain);
  ctx->launches++;
  LG_CUDA(ctx, cudaGetLastError());
  return OK;
}
template <bool DIF, bool COPY0>
static int launch_global_pass_mapped_r(Ctx* ctx, int r, const Fr* in, Fr* out, size_t rows, int q, int s, const FrTw* W,
                                       const OutMap& map, uint32_t rows_per_plane, int copy_plain = 0) {
  if (r == 3) return launch_global_pass_mapped<3, DIF, COPY0>(ctx, in, out, rows, q, s, W, map, rows_per_plane, copy_plain);
  if (r == 2) return launch_global_pass_mapped<2, DIF, COPY0>(ctx, in, out, rows, q, s, W, map, rows_per_plane, copy_plain);
  return launch_global_pass_mapped<1, DIF, COPY0>(ctx, in, out, rows, q, s, W, map, rows_per_plane, copy_plain);
}

int encode_rows(Ctx* ctx, const Fr* msg, size_t rows, int log_k, int rho_inv, Fr* plane0, Fr* cosets, const OutMap* map,
                bool plain_cosets) {
  if (rows == 0) return OK;
  if ((rows << log_k) >= ((size_t)1 << 42) || rows >= ((size_t)1 << 31)) return set_error(ctx, ERR_INVALID, "matrix too large");
  const NttTables* t;
  LG_TRY(get_tables(ctx, log_k, rho_inv, &t, plain_cosets));
  const int q = log_k, l = q < kLogE ? q : kLogE;
  const NttTables* tl = t;
  if (l != q) LG_TRY(get_tables(ctx, l, 1, &tl));
  const size_t total = rows << q;
  LocalArgs a{};
  a.out = cosets;
  a.total = total;
  a.plane_stride = total;
  a.q = q;
  a.l = l;
  a.rho = rho_inv;
  a.w_fwd = tl->w_fwd;
  a.w_inv = tl->w_inv;
  a.scale = t->scale;
  a.final = (q == l) ? 1 : 0;
  a.plain0 = plain_cosets ? 1 : 0;
  if (map) a.map = *map;
  phase_mark(ctx, PH_BEGIN);
  if (q > l) {
    void* tmp;
    LG_TRY(ctx_scratch(ctx, total * sizeof(Fr), &tmp));
    const Fr* src = msg;
    bool first = true;
    for (int top = q; top > l;) {
      const int r = pass_radix(top - l), s = top - r;
      if (first && map)
        LG_TRY((launch_global_pass_mapped_r<true, true>(ctx, r, src, (Fr*)tmp, rows, q, s, t->w_inv, *map, (uint32_t)rows,
                                                        plain_cosets ? 1 : 0)));
      else
        LG_TRY(launch_global_pass_r<true>(ctx, r, src, (Fr*)tmp, first ? plane0 : nullptr, rows, q, s, t->w_inv, 0,
                                          plain_cosets ? 1 : 0));
      src = (const Fr*)tmp;
      first = false;
      top = s;
    }
    a.in = (const Fr*)tmp;
    a.plane0 = nullptr;
    a.mapped = 0;  // the local kernel writes the intermediate; the last strided pass does the mapped stores
    phase_mark(ctx, PH_NTT_STRIDED_INV);
  } else {
    a.in = msg;
    a.plane0 = plane0;
    a.mapped = map ? 1 : 0;
    a.copy0 = map ? 1 : 0;
  }
  if (l == kLogPE && rho_inv > 1 && ctx->persist_groups) {
    a.scale4 = t->scale4;
    LG_TRY(launch_persist(ctx, a, ctx->persist_groups));
  } else {
    LG_TRY(launch_local<0>(ctx, a));
  }
  phase_mark(ctx, PH_NTT_LOCAL);
  if (ctx->ev_after_local) LG_CUDA(ctx, cudaEventRecord(ctx->ev_after_local, ctx->stream));
  if (q > l && rho_inv > 1) {
    Fr* p = cosets;
    const size_t prow = rows * (size_t)(rho_inv - 1);
    for (int s = l; s < q;) {
      const int r = pass_radix(q - s);
      if (map && s + r == q)
        LG_TRY((launch_global_pass_mapped_r<false, false>(ctx, r, p, nullptr, prow, q, s, t->w_fwd, *map, (uint32_t)rows)));
      else
        LG_TRY(launch_global_pass_r<false>(ctx, r, p, p, nullptr, prow, q, s, t->w_fwd, s + r == q ? 1 : 0));
      s += r;
    }
    phase_mark(ctx, PH_NTT_STRIDED_FWD);
  }
  return OK;
}

int intt_rows(Ctx* ctx, const Fr* in, Fr* out, size_t rows, int log_k) {
  if (rows == 0) return OK;
  const NttTables* t;
  LG_TRY(get_tables(ctx, log_k, 1, &t));
  const int q = log_k, l = q < kLogE ? q : kLogE;
  const NttTables* tl = t;
  if (l != q) LG_TRY(get_tables(ctx, l, 1, &tl));
  const size_t total = rows << q;
  LocalArgs a{};
  a.out = out;
  a.total = total;
  a.plane_stride = total;
  a.q = q;
  a.l = l;
  a.rho = 1;
  a.w_fwd = tl->w_fwd;
  a.w_inv = tl->w_inv;
  a.scale = t->scale;
  a.kinv = t->kinv;
  if (q > l) {
    void* tmp;
    LG_TRY(ctx_scratch(ctx, total * sizeof(Fr), &tmp));
    const Fr* src = in;
    for (int top = q; top > l;) {
      const int r = pass_radix(top - l), s = top - r;
      LG_TRY(launch_global_pass_r<true>(ctx, r, src, (Fr*)tmp, nullptr, rows, q, s, t->w_inv));
      src = (const Fr*)tmp;
      top = s;
    }
    a.in = (const Fr*)tmp;
  } else {
    a.in = in;
  }
  return launch_local<1>(ctx, a);
}

}  // namespace lg
