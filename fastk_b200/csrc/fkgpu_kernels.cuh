/*  fkgpu_kernels.cuh -- sm_100a kernels of the FastK counting hot path.
 *
 *  Data model (all integer):
 *    packed reads   seq: 16 bases / u32 word, first base in bits 31:30;  val: 32 positions / u32 word,
 *                   MSB first, 1 = acgt.  Reads are concatenated with one invalid (terminator) position
 *                   between them, so a k-mer window is legal iff its k val bits are all 1
 *                   (restates split.c:1079-1086,1124-1129: short reads and non-acgt windows are dropped).
 *    record         Key<NW>: NW 64-bit words, word 0 most significant, the canonical k-mer left aligned
 *                   (bit 63 of w[0] = high bit of the first base), unused low bits 0.  Integer order of
 *                   (w[0],w[1]) == bytewise order of the reference's KMER_BYTES key (count.c:473-510).
 *
 *  Pipeline (fkgpu_api.cu drives it):
 *    k_scan<HIST>     reads -> histogram of the top P1 key bits               (replaces split.c scan + count.c:1406-1414)
 *    k_scan<SCATTER>  reads -> records grouped by that prefix into buffer A   (replaces Distribute_Block/kmer_list_thread)
 *    k_refine         one CTA per prefix bucket: MSD pass on the next P2 bits, A -> B, cursors in smem (MSDsort.c:129-261)
 *    k_groups         packs consecutive fine buckets into <= C-record work groups
 *    k_sortcount      one CTA per group: TMA bulk load -> smem hash count -> bitonic sort of the distinct keys
 *                     -> run totals, histogram, staged (key,count)            (MSDsort.c:491-509 hist_kmers)
 *    k_compact        staged entries -> [KMER_BYTES key][u16 count] records    (count.c:564-616 table_write_thread)
 */
#pragma once
#ifndef FKGPU_PROBE_CAP
#define FKGPU_PROBE_CAP 64          /* linear-probe steps after which a hash class counts as overflowing */
#endif
#include <cuda_runtime.h>
#include <stdint.h>

namespace fk {

typedef unsigned long long u64;
typedef uint32_t           u32;

template<int NW> struct __align__((NW == 2) ? 16 : 8) Key { u64 w[NW]; };

template<int NW> __device__ __forceinline__ bool key_eq(const Key<NW> &a, const Key<NW> &b)
{ bool e = true;
#pragma unroll
  for (int i = 0; i < NW; i++) e = e && (a.w[i] == b.w[i]);
  return e;
}
template<int NW> __device__ __forceinline__ bool key_lt(const Key<NW> &a, const Key<NW> &b)
{
#pragma unroll
  for (int i = 0; i < NW; i++)
    { if (a.w[i] < b.w[i]) return true;
      if (a.w[i] > b.w[i]) return false;
    }
  return false;
}

/* 64 key bits starting at bit position pos (0 = MSB of w[0]); zero filled past the end */
template<int NW> __device__ __forceinline__ u64 key_bits64(const Key<NW> &a, int pos)
{ if (NW == 1)
    return (pos < 64) ? (a.w[0] << pos) : 0ull;
  else
    { if (pos == 0)  return a.w[0];
      if (pos < 64)  return (a.w[0] << pos) | (a.w[NW > 1 ? 1 : 0] >> (64 - pos));
      if (pos < 128) return a.w[NW > 1 ? 1 : 0] << (pos - 64);
      return 0ull;
    }
}
template<int NW> __device__ __forceinline__ u32 key_digit(const Key<NW> &a, int pos, int nbits)
{ return (u32) (key_bits64<NW>(a,pos) >> (64 - nbits)); }

/* reverse-complement the 16 bases of a packed word */
__device__ __forceinline__ u32 rc32(u32 x)
{ x = __brev(~x);
  return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}

struct V96 { u32 a, b, c; };       /* 96 validity bits, MSB first */

__device__ __forceinline__ V96 shl96(V96 v, int s)
{ V96 r;
  if (s >= 96) { r.a = r.b = r.c = 0; return r; }
  if (s >= 64) { v.a = v.c; v.b = 0; v.c = 0; s -= 64; }
  else if (s >= 32) { v.a = v.b; v.b = v.c; v.c = 0; s -= 32; }
  r.a = __funnelshift_l(v.b,v.a,s);
  r.b = __funnelshift_l(v.c,v.b,s);
  r.c = v.c << s;
  return r;
}
__device__ __forceinline__ V96 and96(V96 x, V96 y) { V96 r; r.a = x.a&y.a; r.b = x.b&y.b; r.c = x.c&y.c; return r; }

/* bit j (MSB first) of the result = 1 iff positions j .. j+k-1 are all valid  (erosion by k) */
__device__ __forceinline__ u32 window_ok(V96 v, int k)
{ V96 acc; acc.a = acc.b = acc.c = 0xffffffffu;
  V96 pw = v;
  int m = 0, n = 1;
  while (n <= k)
    { if (k & n)
        { acc = and96(acc,shl96(pw,m));
          m += n;
        }
      pw = and96(pw,shl96(pw,n));
      n <<= 1;
    }
  return acc.a;
}

/* ---------------------------------------------------------------------------------------------- */

struct ScanParams
  { const u32 *seq;          /* packed bases                                   */
    const u32 *val;          /* validity bits                                  */
    long long  npos;         /* # of positions                                 */
    long long  nseqw;        /* # of real seq words (ceil(npos/16))            */
    long long  nvalw;        /* # of real val words (ceil(npos/32))            */
    int        k;
    int        pbits;        /* P1                                             */
    u32        kmask[4];     /* mask of the 2k key bits over four 32-bit words */
    long long  ntiles, tpc;  /* # of tiles, tiles per CTA                      */
    u32       *cta_hist;     /* HIST out: [grid][2^P1]                         */
    const u64 *cta_off;      /* SCATTER in: [grid][2^P1] offset of the CTA's region inside each bucket */
    const u64 *off1;         /* SCATTER in: bucket starts                      */
    u64       *cursor;       /* scatter variants 1/3: running global cursors (init = bucket starts) */
    void      *out;          /* SCATTER: record buffer                         */
  };

#define SCAN_TPB   256
#define SCAN_PPT   32                       /* k-mer start positions per thread */
#define SCAN_TILE  (SCAN_TPB*SCAN_PPT)
#define SCAN_LHALO 4                        /* seq words kept left of the tile  */
#define SCAN_SEQW  (SCAN_TILE/16 + SCAN_LHALO + 8)
#define SCAN_VALW  (SCAN_TILE/32 + 4)

/*  Per-thread window state: W = the 6 packed words covering the thread's 32 start positions and their
 *  k-1 lookahead, R = the same bases reverse-complemented and reversed, aligned so that the k-mer starting
 *  at local position j is the bit slice [2j, 2j+2k) of W and [62-2j, 62-2j+2k) of R.                      */
struct Window { u32 W[6]; u32 R[6]; u32 ok; };

__device__ __forceinline__ void load_window(Window &w, const u32 *s_seq, const u32 *s_val, int t, int k)
{ const u32 *s = s_seq + SCAN_LHALO + 2*t;
#pragma unroll
  for (int i = 0; i < 6; i++) w.W[i] = s[i];
  const int b  = 2*k + 30;            /* bit offset (from W[0]) of the slice that becomes R[0]   */
  const int q0 = b >> 5, r = b & 31;
#pragma unroll
  for (int i = 0; i < 6; i++)
    w.R[i] = rc32(__funnelshift_l(s[q0-i+1],s[q0-i],r));
  V96 v; v.a = s_val[t]; v.b = s_val[t+1]; v.c = s_val[t+2];
  w.ok = window_ok(v,k);
}

/* canonical key words (32-bit, most significant first) of local k-mer J; NW32 = 2*NW */
template<int NW32, int J>
__device__ __forceinline__ void canon_kmer(const Window &w, const u32 *kmask, u32 *C)
{ u32 F[NW32], G[NW32];
  const int fq = (2*J) >> 5,      fr = (2*J) & 31;
  const int gq = (62-2*J) >> 5,   gr = (62-2*J) & 31;
#pragma unroll
  for (int m = 0; m < NW32; m++)
    { F[m] = __funnelshift_l(w.W[fq+m+1],w.W[fq+m],fr) & kmask[m];
      G[m] = __funnelshift_l(w.R[gq+m+1],w.R[gq+m],gr) & kmask[m];
    }
  bool lt = false, dec = false;      /* lt = (G < F) decided at the first differing word */
#pragma unroll
  for (int m = 0; m < NW32; m++)
    { if (!dec && F[m] != G[m]) { lt = (G[m] < F[m]); dec = true; } }
#pragma unroll
  for (int m = 0; m < NW32; m++) C[m] = lt ? G[m] : F[m];
}

template<int NW32, int J> struct KmerLoop
{ template<class Fn> __device__ __forceinline__ static void run(const Window &w, const u32 *kmask, Fn &fn)
  { if ((w.ok >> (31-J)) & 1u)
      { u32 C[NW32];
        canon_kmer<NW32,J>(w,kmask,C);
        fn(J,C);
      }
    KmerLoop<NW32,J+1>::run(w,kmask,fn);
  }
};
template<int NW32> struct KmerLoop<NW32,SCAN_PPT>
{ template<class Fn> __device__ __forceinline__ static void run(const Window &, const u32 *, Fn &) {} };

__device__ __forceinline__ void scan_load_tile(const ScanParams &p, long long tile, u32 *s_seq, u32 *s_val)
{ const long long w0 = tile * (SCAN_TILE/16) - SCAN_LHALO;
  for (int i = threadIdx.x; i < SCAN_SEQW; i += SCAN_TPB)
    { long long g = w0 + i;
      s_seq[i] = (g >= 0 && g < p.nseqw) ? __ldg(p.seq+g) : 0u;
    }
  const long long v0 = tile * (SCAN_TILE/32);
  for (int i = threadIdx.x; i < SCAN_VALW; i += SCAN_TPB)
    { long long g = v0 + i;
      s_val[i] = (g < p.nvalw) ? __ldg(p.val+g) : 0u;
    }
}

/*  Persistent: CTA c owns the contiguous tile range [c*tpc, (c+1)*tpc).  HIST accumulates its private
 *  histogram in smem and stores it (no atomics); k_colscan turns the per-CTA histograms into per-CTA output
 *  regions inside every bucket; SCATTER then needs only smem cursors: one pass, no global atomics.          */
template<int NW, bool SCATTER>
__global__ void __launch_bounds__(SCAN_TPB) k_scan(ScanParams p)
{ extern __shared__ u32 s_dyn[];
  u32 *s_seq  = s_dyn;
  u32 *s_val  = s_seq + SCAN_SEQW;
  u32 *s_hist = s_val + SCAN_VALW;                 /* [2^P1] counts                    */
  const int nb = 1 << p.pbits;
  u64 *s_base = (u64 *) (s_hist + nb + (nb & 1));  /* [2^P1] region starts, SCATTER only */
  constexpr int NW32 = 2*NW;
  const int sh = 32 - p.pbits;
  const long long t0 = (long long) blockIdx.x * p.tpc;
  long long t1 = t0 + p.tpc; if (t1 > p.ntiles) t1 = p.ntiles;

  for (int i = threadIdx.x; i < nb; i += SCAN_TPB)
    { s_hist[i] = 0;
      if (SCATTER) s_base[i] = p.off1[i] + p.cta_off[(u64) blockIdx.x * nb + i];
    }
  u32 km[4] = { p.kmask[0], p.kmask[1], p.kmask[2], p.kmask[3] };
  Key<NW> *out = (Key<NW> *) p.out;

  for (long long t = t0; t < t1; t++)
    { __syncthreads();
      scan_load_tile(p,t,s_seq,s_val);
      __syncthreads();
      Window w;
      load_window(w,s_seq,s_val,threadIdx.x,p.k);
      if (!SCATTER)
        { auto fn = [&](int, const u32 *C) { atomicAdd(&s_hist[p.pbits ? (C[0] >> sh) : 0u],1u); };
          KmerLoop<NW32,0>::run(w,km,fn);
        }
      else
        { auto fb = [&](int, const u32 *C)
            { u32 d = p.pbits ? (C[0] >> sh) : 0u;
              u32 r = atomicAdd(&s_hist[d],1u);
              Key<NW> key;
#pragma unroll
              for (int m = 0; m < NW; m++) key.w[m] = ((u64) C[2*m] << 32) | C[2*m+1];
              out[s_base[d] + r] = key;
            };
          KmerLoop<NW32,0>::run(w,km,fb);
        }
    }
  if (!SCATTER)
    { __syncthreads();
      for (int i = threadIdx.x; i < nb; i += SCAN_TPB)
        p.cta_hist[(u64) blockIdx.x * nb + i] = s_hist[i];
    }
}

/*  Scatter variant 1: one tile per CTA, two compute passes; pass A ranks the tile's k-mers in smem, one global
 *  atomic per (tile, non-empty bucket) reserves a run at the bucket's shared frontier, pass B writes.          */
template<int NW>
__global__ void __launch_bounds__(SCAN_TPB) k_scatter_tile(ScanParams p)
{ extern __shared__ u32 s_dyn[];
  u32 *s_seq  = s_dyn;
  u32 *s_val  = s_seq + SCAN_SEQW;
  u32 *s_hist = s_val + SCAN_VALW;
  const int nb = 1 << p.pbits;
  u64 *s_base = (u64 *) (s_hist + nb + (nb & 1));
  constexpr int NW32 = 2*NW;
  const int sh = 32 - p.pbits;
  for (int i = threadIdx.x; i < nb; i += SCAN_TPB) s_hist[i] = 0;
  scan_load_tile(p,blockIdx.x,s_seq,s_val);
  __syncthreads();
  Window w;
  load_window(w,s_seq,s_val,threadIdx.x,p.k);
  u32 km[4] = { p.kmask[0], p.kmask[1], p.kmask[2], p.kmask[3] };
  u32 rk[SCAN_PPT/2];
#pragma unroll
  for (int i = 0; i < SCAN_PPT/2; i++) rk[i] = 0;
  auto fa = [&](int j, const u32 *C)
    { u32 r = atomicAdd(&s_hist[p.pbits ? (C[0] >> sh) : 0u],1u);
      rk[j>>1] |= r << (16*(j&1));
    };
  KmerLoop<NW32,0>::run(w,km,fa);
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += SCAN_TPB)
    { u32 c = s_hist[i];
      s_base[i] = c ? atomicAdd(p.cursor+i,(u64) c) : 0ull;
    }
  __syncthreads();
  Key<NW> *out = (Key<NW> *) p.out;
  auto fb = [&](int j, const u32 *C)
    { u32 d = p.pbits ? (C[0] >> sh) : 0u;
      u32 r = (rk[j>>1] >> (16*(j&1))) & 0xffffu;
      Key<NW> key;
#pragma unroll
      for (int m = 0; m < NW; m++) key.w[m] = ((u64) C[2*m] << 32) | C[2*m+1];
      out[s_base[d] + r] = key;
    };
  KmerLoop<NW32,0>::run(w,km,fb);
}

/*  Scatter variant 3: single pass, one global atomic (with return) per k-mer on the bucket's shared frontier. */
template<int NW>
__global__ void __launch_bounds__(SCAN_TPB) k_scatter_atomic(ScanParams p)
{ extern __shared__ u32 s_dyn[];
  u32 *s_seq  = s_dyn;
  u32 *s_val  = s_seq + SCAN_SEQW;
  constexpr int NW32 = 2*NW;
  const int sh = 32 - p.pbits;
  scan_load_tile(p,blockIdx.x,s_seq,s_val);
  __syncthreads();
  Window w;
  load_window(w,s_seq,s_val,threadIdx.x,p.k);
  u32 km[4] = { p.kmask[0], p.kmask[1], p.kmask[2], p.kmask[3] };
  Key<NW> *out = (Key<NW> *) p.out;
  auto fb = [&](int, const u32 *C)
    { u32 d = p.pbits ? (C[0] >> sh) : 0u;
      u64 ps = atomicAdd(p.cursor + d,1ull);
      Key<NW> key;
#pragma unroll
      for (int m = 0; m < NW; m++) key.w[m] = ((u64) C[2*m] << 32) | C[2*m+1];
      out[ps] = key;
    };
  KmerLoop<NW32,0>::run(w,km,fb);
}

/*  per bucket d: total[d] = sum_c cta_hist[c][d];  cta_off[c][d] = sum_{c' < c} cta_hist[c'][d]             */
__global__ void k_colscan(const u32 *cta_hist, u64 *cta_off, u64 *total, int ncta, int nb)
{ int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nb) return;
  u64 run = 0;
  for (int c = 0; c < ncta; c++)
    { u32 x = cta_hist[(u64) c * nb + d];
      cta_off[(u64) c * nb + d] = run;
      run += x;
    }
  total[d] = run;
}

/* ---------------------------------------------------------------------------------------------- */
/*  ASCII -> packed.  One thread = 32 positions = 2 seq words + 1 val word.                          */

__global__ void k_pack_ascii(const uint4 *ascii, long long npos, u32 *seq, u32 *val, long long nvalw)
{ long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nvalw) return;
  u32 s0 = 0, s1 = 0, v = 0;
  long long base = t*32;
#pragma unroll
  for (int q = 0; q < 2; q++)
    { uint4 x = make_uint4(0,0,0,0);
      if (base + q*16 < npos) x = __ldg(ascii + t*2 + q);       /* buffer is padded to 32 bytes */
      u32 wd[4] = { x.x, x.y, x.z, x.w };
#pragma unroll
      for (int i = 0; i < 16; i++)
        { u32 c = (wd[i>>2] >> (8*(i&3))) & 0xffu;
          u32 lc = c | 0x20u;
          bool ok = (lc == 'a') | (lc == 'c') | (lc == 'g') | (lc == 't');
          ok = ok && (base + q*16 + i < npos);
          u32 code = ((c >> 1) ^ (c >> 2)) & 3u;
          code = ok ? code : 0u;
          if (q == 0) s0 |= code << (30 - 2*i); else s1 |= code << (30 - 2*i);
          v |= (ok ? 1u : 0u) << (31 - (q*16 + i));
        }
    }
  seq[2*t] = s0; seq[2*t+1] = s1; val[t] = v;
}

/*  -bc<n>: invalidate the first bc positions of every read (split.c:1075 s += BC_PREFIX).          */
__global__ void k_mask_prefix(const long long *rstart, long long nreads, int bc, long long npos, u32 *val)
{ long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nreads) return;
  long long p = rstart[r];
  for (int i = 0; i < bc && p + i < npos; i++)
    atomicAnd(val + ((p+i) >> 5), ~(1u << (31 - ((p+i) & 31))));
}

/* ---------------------------------------------------------------------------------------------- */
/*  Small utility kernels                                                                           */

/* exclusive scan of a u64 array of n <= 8192 entries, one CTA; out[n] = total; optional copy to cur */
__global__ void k_scan_small(const u64 *in, u64 *out, u64 *cur, int n)
{ __shared__ u64 s_part[1024];
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int b = threadIdx.x * per;
  u64 sum = 0;
  for (int i = b; i < b+per && i < n; i++) sum += in[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0)
    { u64 run = 0;
      for (int i = 0; i < (int) blockDim.x; i++) { u64 x = s_part[i]; s_part[i] = run; run += x; }
    }
  __syncthreads();
  u64 run = s_part[threadIdx.x];
  for (int i = b; i < b+per && i < n; i++)
    { u64 x = in[i]; out[i] = run; if (cur) cur[i] = run; run += x; }
  if (threadIdx.x == blockDim.x-1) out[n] = run;
}

/* three-phase exclusive scan of u32 -> u64 for large n (chunk = 2048 per CTA, 256 threads x 8) */
#define LS_CHUNK 2048
__global__ void __launch_bounds__(256) k_lscan_reduce(const u32 *in, long long n, u64 *bsum)
{ __shared__ u64 s_w[8];
  long long base = (long long) blockIdx.x * LS_CHUNK;
  u64 sum = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
    { long long g = base + threadIdx.x*8 + i;
      if (g < n) sum += in[g];
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu,sum,o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0)
    { u64 t = 0;
      for (int i = 0; i < 8; i++) t += s_w[i];
      bsum[blockIdx.x] = t;
    }
}
__global__ void k_lscan_top(u64 *bsum, long long nb, u64 *total)
{ /* single CTA, sequential over chunks of blockDim */
  __shared__ u64 s_part[1024];
  __shared__ u64 s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (long long b0 = 0; b0 < nb; b0 += blockDim.x)
    { long long g = b0 + threadIdx.x;
      u64 x = (g < nb) ? bsum[g] : 0;
      s_part[threadIdx.x] = x;
      __syncthreads();
      /* Hillis-Steele inclusive scan */
      for (int o = 1; o < (int) blockDim.x; o <<= 1)
        { u64 y = (threadIdx.x >= (unsigned) o) ? s_part[threadIdx.x-o] : 0;
          __syncthreads();
          s_part[threadIdx.x] += y;
          __syncthreads();
        }
      u64 incl = s_part[threadIdx.x];
      u64 carry = s_carry;
      if (g < nb) bsum[g] = carry + incl - x;
      __syncthreads();
      if (threadIdx.x == blockDim.x-1) s_carry = carry + incl;
      __syncthreads();
    }
  if (threadIdx.x == 0) *total = s_carry;
}
__global__ void __launch_bounds__(256) k_lscan_apply(const u32 *in, long long n, const u64 *bsum, u64 *out)
{ __shared__ u64 s_w[8];
  long long base = (long long) blockIdx.x * LS_CHUNK;
  u32 x[8];
  u64 sum = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
    { long long g = base + threadIdx.x*8 + i;
      x[i] = (g < n) ? in[g] : 0;
      sum += x[i];
    }
  u64 incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
    { u64 y = __shfl_up_sync(0xffffffffu,incl,o);
      if ((threadIdx.x & 31) >= o) incl += y;
    }
  if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
  __syncthreads();
  u64 woff = 0;
  for (int i = 0; i < (int) (threadIdx.x >> 5); i++) woff += s_w[i];
  u64 run = bsum[blockIdx.x] + woff + incl - sum;
#pragma unroll
  for (int i = 0; i < 8; i++)
    { long long g = base + threadIdx.x*8 + i;
      if (g < n) out[g] = run;
      run += x[i];
    }
}

/* ---------------------------------------------------------------------------------------------- */
/*  k_refine: one CTA per level-1 bucket (persistent, ticketed).  MSD pass on bits [pos, pos+nbits):
 *  histogram in smem, block scan, scatter with smem cursors (no global atomics).  Writes the child
 *  start offsets off2[(b << nbits) + d].                                                            */

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar)
{ asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{ asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "WAIT_LOOP:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
    "@p bra WAIT_DONE;\n\t"
    "bra WAIT_LOOP;\n\t"
    "WAIT_DONE:\n\t}"
    :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}


#define REF_TPB   256
#define REF_TB    1024                     /* records per smem stage                 */
#define REF_ST    4                        /* stages in flight (TMA bulk copies)      */
#define REF_SLOT  (REF_TB + 2)             /* slot stride in records (8-byte keys may start odd) */

template<int NW>
__global__ void __launch_bounds__(REF_TPB) k_refine(const Key<NW> *__restrict__ src, Key<NW> *__restrict__ dst,
                                                    const u64 *off1, int nb1, int pos, int nbits,
                                                    u64 *off2, u32 *ticket)
{ extern __shared__ __align__(128) unsigned char s_raw[];
  Key<NW> *sbuf = (Key<NW> *) s_raw;
  u32 *s_cnt = (u32 *) (s_raw + (size_t) REF_ST * REF_SLOT * sizeof(Key<NW>));     /* [2^nbits] */
  __shared__ u64 s_full[REF_ST];
  __shared__ u32 s_warp[REF_TPB/32];
  __shared__ int s_b;
  const int nd = 1 << nbits;
  u32 q = 0;                            /* running tile sequence number of this CTA (slot = q % ST, parity = q / ST) */

  if (threadIdx.x == 0)
    for (int i = 0; i < REF_ST; i++) mbar_init(&s_full[i],1);
  __syncthreads();

  for (;;)
    { if (threadIdx.x == 0) s_b = (int) atomicAdd(ticket,1u);
      __syncthreads();
      const int b = s_b;
      if (b >= nb1) break;
      const u64 start = off1[b];
      const u64 n = off1[b+1] - start;
      const u32 shift = (NW & 1) ? (u32) (start & 1ull) : 0u;
      const u64 ntl = (n + REF_TB - 1) / REF_TB;
      for (int i = threadIdx.x; i < nd; i += REF_TPB) s_cnt[i] = 0;
      __syncthreads();

      for (int pass = 0; pass < 2; pass++)
        { auto issue = [&](u64 t, u32 qq)
            { u64 r0 = start + t * REF_TB;
              u64 cnt = n - t * REF_TB; if (cnt > REF_TB) cnt = REF_TB;
              u64 a0 = r0 - shift;
              u64 len = cnt + shift; if (NW & 1) len = (len + 1) & ~1ull;
              u32 bytes = (u32) (len * sizeof(Key<NW>));
              u32 sl = qq % REF_ST;
              mbar_expect_tx(&s_full[sl],bytes);
              tma_bulk_g2s(sbuf + (size_t) sl * REF_SLOT,src + a0,bytes,&s_full[sl]);
            };
          if (threadIdx.x == 0)
            for (u64 t = 0; t < ntl && t < REF_ST; t++) issue(t,q + (u32) t);
          for (u64 t = 0; t < ntl; t++)
            { const u32 sl = q % REF_ST;
              mbar_wait(&s_full[sl],(q / REF_ST) & 1u);
              const Key<NW> *tile = sbuf + (size_t) sl * REF_SLOT + shift;
              u64 cnt = n - t * REF_TB; if (cnt > REF_TB) cnt = REF_TB;
              if (pass == 0)
                {
#pragma unroll
                  for (int u = 0; u < REF_TB/REF_TPB; u++)
                    { u32 i = u*REF_TPB + threadIdx.x;
                      if (i < cnt) atomicAdd(&s_cnt[key_digit<NW>(tile[i],pos,nbits)],1u);
                    }
                }
              else
                { Key<NW> *out = dst + start;
#pragma unroll
                  for (int u = 0; u < REF_TB/REF_TPB; u++)
                    { u32 i = u*REF_TPB + threadIdx.x;
                      if (i < cnt)
                        { Key<NW> r = tile[i];
                          u32 ps = atomicAdd(&s_cnt[key_digit<NW>(r,pos,nbits)],1u);
                          out[ps] = r;
                        }
                    }
                }
              __syncthreads();
              if (threadIdx.x == 0 && t + REF_ST < ntl) issue(t + REF_ST,q + REF_ST);
              q++;
            }
          if (pass == 0)
            { /* block exclusive scan of s_cnt; publish the child starts */
              const int per = (nd + REF_TPB - 1) / REF_TPB;
              const int b0 = threadIdx.x * per;
              u32 sum = 0;
              for (int i = b0; i < b0+per && i < nd; i++) sum += s_cnt[i];
              u32 incl = sum;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1)
                { u32 y = __shfl_up_sync(0xffffffffu,incl,o);
                  if ((threadIdx.x & 31) >= o) incl += y;
                }
              if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
              __syncthreads();
              u32 woff = 0;
              for (int i = 0; i < (int) (threadIdx.x >> 5); i++) woff += s_warp[i];
              u32 run = woff + incl - sum;
              for (int i = b0; i < b0+per && i < nd; i++)
                { u32 c = s_cnt[i];
                  s_cnt[i] = run;
                  off2[((u64) b << nbits) + i] = start + run;
                  run += c;
                }
              __syncthreads();
            }
        }
      if (b == nb1-1 && threadIdx.x == 0) off2[(u64) nb1 << nbits] = off1[nb1];
      __syncthreads();
    }
}

/* ---------------------------------------------------------------------------------------------- */
/*  Tile-parallel histogram / partition of ONE flat record array on its top `nbits` bits (level 1 when
 *  the input is records rather than reads: the multi-GPU path after the exchange).                   */

#define TP_TPB 256
#define TP_RPT(NW) ((NW) == 3 ? 24 : 32)
#define TP_TILE(NW) (TP_TPB*TP_RPT(NW))

/*  One CTA per tile of TP_TILE records.  The tile is read twice -- once to rank every record inside its bucket (smem atomics;
 *  only the 16-bit ranks stay in registers), once more (an L2 hit) to scatter -- so that a tile can be 8192 records without
 *  holding them in registers: the global atomics that reserve a run per (tile, non-empty bucket) were what this kernel
 *  waited for (ncu r2: long_scoreboard 22 of 47 stalled warps), and their number per record halves with the tile size.   */
template<int NW, bool SCATTER>
__global__ void __launch_bounds__(TP_TPB) k_tilepart(const Key<NW> *__restrict__ src, Key<NW> *__restrict__ dst,
                                                     u64 n, int nbits, u64 *hist)
{ extern __shared__ u32 s_dyn[];
  u32 *s_cnt = s_dyn;
  const int nd = 1 << nbits;
  u64 *s_base = (u64 *) (s_cnt + nd + (nd & 1));
  for (int i = threadIdx.x; i < nd; i += TP_TPB) s_cnt[i] = 0;
  __syncthreads();
  constexpr int RPT = TP_RPT(NW);
  const u64 t0 = (u64) blockIdx.x * TP_TILE(NW);
  u32 rk[RPT/2];
#pragma unroll
  for (int u = 0; u < RPT/2; u++) rk[u] = 0;
#pragma unroll
  for (int u = 0; u < RPT; u++)
    { u64 i = t0 + u*TP_TPB + threadIdx.x;
      if (i < n)
        { const Key<NW> r = src[i];
          u32 rr = atomicAdd(&s_cnt[nbits ? key_digit<NW>(r,0,nbits) : 0u],1u);
          rk[u>>1] |= rr << (16*(u&1));
        }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < nd; i += TP_TPB)
    { u32 c = s_cnt[i];
      if (SCATTER) s_base[i] = c ? atomicAdd(hist+i,(u64) c) : 0ull;
      else if (c) atomicAdd(hist+i,(u64) c);
    }
  if (!SCATTER) return;
  __syncthreads();
#pragma unroll
  for (int u = 0; u < RPT; u++)
    { u64 i = t0 + u*TP_TPB + threadIdx.x;
      if (i < n)
        { const Key<NW> r = src[i];
          u32 d = nbits ? key_digit<NW>(r,0,nbits) : 0u;
          dst[s_base[d] + ((rk[u>>1] >> (16*(u&1))) & 0xffffu)] = r;
        }
    }
}

/* ---------------------------------------------------------------------------------------------- */
/*  k_groups: gstart[g] = smallest fine-bucket start >= g*T  (so group g = [gstart[g], gstart[g+1]))  */

__global__ void k_fill_u64(u64 *a, long long n, const u64 *value)
{ long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = *value;
}

__global__ void k_groups(const u64 *off, long long m /* # of buckets; off[m] = end */, u32 T, u64 *gstart, long long gmax)
{ long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i > m) return;
  const u64 base = off[0];                   /* a bucket range need not start at record 0 */
  u64 s = off[i];
  long long glo = (i == 0) ? 0 : (long long) ((off[i-1] - base) / T) + 1;
  long long ghi = (long long) ((s - base) / T);
  if (i == 0) glo = 0;
  for (long long g = glo; g <= ghi && g <= gmax; g++)
    gstart[g] = s;
}

/* ---------------------------------------------------------------------------------------------- */
/*  k_sortcount: one CTA per work item [start, end) of <= cap records.                               */

#define SC_TPB 256
#define SC_EMPTY 0xffffffffu
#define SC_PAD   0xffffffffffffffffull
#define SC_SMALLHIST 256
#define SC_BINBITS   10
#define SC_NBIN      (1 << SC_BINBITS)

#define ITEM_UNIFORM 1u      /* every record of the item holds the same key                  */
#define ITEM_ALTBUF  2u      /* item lives in the alternate buffer (input/staging swapped)    */

struct SortCountParams
  { const void *in0;  void *stage0;      /* default: read in0, stage into stage0              */
    const void *in1;  void *stage1;      /* ITEM_ALTBUF: read in1, stage into stage1          */
    u32        *stage_cnt;               /* count of staged entry, indexed like the stage     */
    const u64  *starts; const u64 *ends; /* item record ranges                                */
    const u32  *flags;                   /* per item, may be NULL                             */
    u32        *e_all;                   /* out: # distinct keys of the item                  */
    u32        *e_pass;                  /* out: # of them with count >= cutoff               */
    u64        *g_hist;                  /* [32768]                                           */
    u64        *g_maxinst;
    u64        *g_ndistinct;
    u32        *ovf_cnt; u32 *ovf_list; u32 ovf_cap;
    u32         item_base;               /* index of item 0 of this launch in the host's item numbering (chunked launches) */
    u32         cap;                     /* C                                                 */
    u32         tab_off, srt_off;        /* byte offsets of the hash table / sort array in smem */
    u32         srt2_off;                /* entries: second sort array inside srt (>= SC_RANKMAX free) */
    u32         weighted;                /* 1: records are distinct (key | count in the low 16 bits): sort only */
    u32         cutoff;
    long long   nitems;
    uint8_t    *direct;                  /* weighted only: write the final [kbytes key][u16 LE count] table records of entry
                                            r0 + q at direct + (r0 + q) * (kbytes + 2) (every entry passes the cutoff)      */
    int         kbytes;
    u32         tab_bytes;               /* bytes of the hash-table region (staging of the direct output)                   */
  };

template<int NW> __device__ __forceinline__ u32 key_hash(const Key<NW> &a)
{ u64 h = a.w[0] * 0x9E3779B97F4A7C15ull;
  if (NW > 1) h ^= a.w[NW > 1 ? 1 : 0] * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  return (u32) (h >> 32);
}

template<int NW>
__global__ void __launch_bounds__(SC_TPB) k_sortcount(SortCountParams p)
{ extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ u64 s_bar;
  __shared__ u32 s_D, s_pass;
  __shared__ u64 s_mn, s_mx;
  __shared__ u32 s_hist[SC_SMALLHIST];
  __shared__ u32 s_bin[SC_NBIN], s_boff[SC_NBIN+1], s_wtot[SC_TPB/32];

  const long long g = blockIdx.x;
  if (g >= p.nitems) return;
  const u64 r0 = p.starts[g], r1 = p.ends[g];
  const u64 n64 = r1 - r0;
  const u32 fl = p.flags ? p.flags[g] : 0u;
  const Key<NW> *in    = (const Key<NW> *) ((fl & ITEM_ALTBUF) ? p.in1 : p.in0);
  Key<NW>       *stage = (Key<NW> *) ((fl & ITEM_ALTBUF) ? p.stage1 : p.stage0);

  if (n64 == 0)
    { if (threadIdx.x == 0) { p.e_all[g] = 0; p.e_pass[g] = 0; }
      return;
    }
  if (fl & ITEM_UNIFORM)
    { /* all n records equal: one entry, count n (saturating; MSDsort.c:498-504) */
      if (threadIdx.x == 0)
        { u32 c = (n64 >= 0x7fffull) ? 0x7fffu : (u32) n64;
          stage[r0] = in[r0];
          p.stage_cnt[r0] = c;
          atomicAdd(p.g_hist + c,1ull);
          if (n64 >= 0x7fffull) atomicAdd(p.g_maxinst,n64);
          atomicAdd(p.g_ndistinct,1ull);
          p.e_all[g] = 1;
          p.e_pass[g] = (c >= p.cutoff) ? 1u : 0u;
        }
      return;
    }
  if (n64 > p.cap)
    { if (threadIdx.x == 0)
        { u32 s = atomicAdd(p.ovf_cnt,1u);
          if (s < p.ovf_cap) p.ovf_list[s] = p.item_base + (u32) g;
          p.e_all[g] = 0; p.e_pass[g] = 0;
        }
      return;
    }
  const u32 n = (u32) n64;

  /* smem carve-up: records | hash table (2*pow2) | sort keys */
  Key<NW> *rec   = (Key<NW> *) s_raw;
  u32      H     = 128; while (H < n + (n >> 2) + 1) H <<= 1;     /* load <= 0.8 even if all keys differ */
  u32     *table = (u32 *) (s_raw + p.tab_off);
  u64     *srt   = (u64 *) (s_raw + p.srt_off);

  /* TMA bulk load of the item (16-byte granules) */
  const u64 a0 = (NW & 1) ? (r0 & ~1ull) : r0;
  const u64 a1 = (NW & 1) ? ((r1 + 1) & ~1ull) : r1;
  const u32 shift = (u32) (r0 - a0);
  const u32 bytes = (u32) ((a1 - a0) * sizeof(Key<NW>));
  if (threadIdx.x == 0)
    { mbar_init(&s_bar,1);
      s_D = 0; s_pass = 0;
    }
  if (threadIdx.x == 0) { s_mn = ~0ull; s_mx = 0ull; }
  for (u32 i = threadIdx.x; i < SC_SMALLHIST; i += SC_TPB) s_hist[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0)
    { mbar_expect_tx(&s_bar,bytes);
      tma_bulk_g2s(rec,in + a0,bytes,&s_bar);
    }
  for (u32 i = threadIdx.x; i < H; i += SC_TPB) table[i] = SC_EMPTY;
  mbar_wait(&s_bar,0);
  __syncthreads();
  rec += shift;

  /* the group's keys span [mn, mx] in their top 64 bits: they are ordered by the top 32 bits of (w0 - mn) scaled to that span,
     so the bins below fill evenly whatever power-of-two boundary the key range straddles (ties fall back to a full compare) */
  u64 mn = ~0ull, mx = 0ull;
  if (p.weighted)
    { /* every record is already a distinct key carrying its count: nothing to merge, only to order */
      for (u32 i = threadIdx.x; i < n; i += SC_TPB)
        { const u64 w0 = rec[i].w[0];
          mn = w0 < mn ? w0 : mn; mx = w0 > mx ? w0 : mx;
        }
    }
  else
  /* hash count: table slot = (owner record index << 16) | multiplicity */
  { for (u32 i = threadIdx.x; i < n; i += SC_TPB)
      { const Key<NW> key = rec[i];
        mn = key.w[0] < mn ? key.w[0] : mn; mx = key.w[0] > mx ? key.w[0] : mx;
        u32 h = key_hash<NW>(key) & (H-1);
        for (;;)
          { u32 cur = ((volatile u32 *) table)[h];
            if (cur == SC_EMPTY)
              { u32 old = atomicCAS(&table[h],SC_EMPTY,(i << 16) | 1u);
                if (old == SC_EMPTY) break;
                cur = old;
              }
            if (key_eq<NW>(rec[cur >> 16],key))
              { atomicAdd(&table[h],1u);
                break;
              }
            h = (h+1) & (H-1);
          }
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    { const u64 a = __shfl_xor_sync(0xffffffffu,mn,o), b = __shfl_xor_sync(0xffffffffu,mx,o);
      mn = a < mn ? a : mn; mx = b > mx ? b : mx;
    }
  if ((threadIdx.x & 31) == 0) { atomicMin(&s_mn,mn); atomicMax(&s_mx,mx); }
  __syncthreads();
  mn = s_mn; mx = s_mx;
  const int nsh = (mx > mn) ? __clzll((long long) (mx - mn)) : 0;
#define SC_NORM32(w0) ((mx > mn) ? ((((w0) - mn) << nsh) >> 32) : 0ull)

  /* gather the distinct keys */
  if (p.weighted)
    { for (u32 i = threadIdx.x; i < n; i += SC_TPB)
        srt[i] = (SC_NORM32(rec[i].w[0]) << 32) | ((u64) i << 16);
      if (threadIdx.x == 0) s_D = n;
    }
  else
  for (u32 s = threadIdx.x; s < H; s += SC_TPB)
    { u32 v = table[s];
      if (v != SC_EMPTY)
        { u32 ps = atomicAdd(&s_D,1u);
          srt[ps] = (SC_NORM32(rec[v >> 16].w[0]) << 32) | v;
        }
    }
  __syncthreads();
  const u32 D = s_D;

  /* order the D distinct keys: bin on the 10 bits that follow the common prefix, then rank inside the (tiny) bins
     by counting -- 4 barriers, no sorting network (the bitonic sort's barriers were 29 % of this kernel's time)   */
  u64 *srtB = (u64 *) table;                       /* the hash table is dead now; it holds >= D entries of 8 bytes   */
  for (u32 i = threadIdx.x; i < SC_NBIN; i += SC_TPB) s_bin[i] = 0;
  __syncthreads();
  for (u32 e = threadIdx.x; e < D; e += SC_TPB)
    atomicAdd(&s_bin[(u32) (srt[e] >> (64 - SC_BINBITS))],1u);
  __syncthreads();
  { /* exclusive scan of the SC_NBIN counters: SC_NBIN / SC_TPB per thread + warp scan + warp totals */
    constexpr int PER = SC_NBIN / SC_TPB;
    u32 v[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; i++) { v[i] = s_bin[threadIdx.x*PER + i]; sum += v[i]; }
    u32 incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
      { u32 y = __shfl_up_sync(0xffffffffu,incl,o);
        if ((threadIdx.x & 31) >= o) incl += y;
      }
    if ((threadIdx.x & 31) == 31) s_wtot[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 woff = 0;
    for (int i = 0; i < (int) (threadIdx.x >> 5); i++) woff += s_wtot[i];
    u32 run = woff + incl - sum;
#pragma unroll
    for (int i = 0; i < PER; i++)
      { s_boff[threadIdx.x*PER + i] = run; s_bin[threadIdx.x*PER + i] = run; run += v[i]; }
    if (threadIdx.x == SC_TPB-1) s_boff[SC_NBIN] = run;
  }
  __syncthreads();
  for (u32 e = threadIdx.x; e < D; e += SC_TPB)
    { const u64 a = srt[e];
      srtB[atomicAdd(&s_bin[(u32) (a >> (64 - SC_BINBITS))],1u)] = a;
    }
  __syncthreads();
  for (u32 e = threadIdx.x; e < D; e += SC_TPB)
    { const u64 a = srtB[e];
      const u32 bn = (u32) (a >> (64 - SC_BINBITS));
      const u32 lo = s_boff[bn], hi = s_boff[bn+1];
      const u32 pa = (u32) (a >> 32);
      u32 rank = lo;
      for (u32 j = lo; j < hi; j++)
        { const u64 b = srtB[j];
          const u32 pb = (u32) (b >> 32);
          bool lt = pb < pa;
          if (pb == pa && j != e)
            { /* entries that agree in every bit (the same k-mer with the same count from two merged tables) order by position */
              const Key<NW> kb = rec[(u32) b >> 16], ka = rec[(u32) a >> 16];
              lt = key_lt<NW>(kb,ka) || (j < e && key_eq<NW>(kb,ka));
            }
          rank += lt ? 1u : 0u;
        }
      srt[rank] = a;
    }
  const u64 *fin = srt;
  __syncthreads();

  if (p.weighted && p.direct != NULL)
    { /* the entries are distinct, all of them pass the cutoff and their final positions are known: write the table records
         straight out (staged through the dead hash-table region so that the global stores are whole words)              */
      const int tw = p.kbytes + 2;
      uint8_t *gout = p.direct + r0 * (u64) tw;
      const u32 a0 = (u32) ((uintptr_t) gout & 3u);
      const u32 total = a0 + D * (u32) tw;
      uint8_t *sout = (uint8_t *) table;
      const bool staged = (total + 4 <= p.tab_bytes);
      for (u32 q = threadIdx.x; q < D; q += SC_TPB)
        { const u32 v = (u32) fin[q];
          Key<NW> key = rec[v >> 16];
          const u32 c = (u32) (key.w[NW-1] & 0xffffull);
          key.w[NW-1] &= ~0xffffull;
          if (staged && tw == 12 && a0 == 0 && NW == 2)
            { u32 *o = (u32 *) sout + q*3;
              o[0] = __byte_perm((u32) (key.w[0] >> 32),0,0x0123);
              o[1] = __byte_perm((u32) key.w[0],0,0x0123);
              o[2] = (u32) (key.w[NW-1] >> 56) | (((u32) (key.w[NW-1] >> 48) & 0xffu) << 8) | (c << 16);
            }
          else
            { uint8_t *e = staged ? (sout + a0 + q * (u32) tw) : (gout + q * (u64) tw);
              for (int b = 0; b < p.kbytes; b++)
                e[b] = (uint8_t) (key.w[b >> 3] >> (56 - 8*(b & 7)));
              e[p.kbytes]   = (uint8_t) (c & 0xffu);
              e[p.kbytes+1] = (uint8_t) (c >> 8);
            }
        }
      if (staged)
        { __syncthreads();
          u32 *gw = (u32 *) (gout - a0);
          const u32 *sw = (const u32 *) sout;
          for (u32 i = threadIdx.x; 4*i < total; i += SC_TPB)
            { if (4*i >= a0 && 4*i + 4 <= total) gw[i] = sw[i];
              else
                for (u32 b = 4*i; b < 4*i + 4; b++)
                  if (b >= a0 && b < total) ((uint8_t *) gw)[b] = sout[b];
            }
        }
      if (threadIdx.x == 0) { p.e_all[g] = D; p.e_pass[g] = D; }
      return;
    }

  /* emit: staged (key,count), histogram */
  u32 npass = 0;
  for (u32 q = threadIdx.x; q < D; q += SC_TPB)
    { u32 v = (u32) fin[q];
      u32 c = v & 0xffffu;
      Key<NW> key = rec[v >> 16];
      if (p.weighted)
        { c = (u32) (key.w[NW-1] & 0xffffull);
          key.w[NW-1] &= ~0xffffull;
        }
      stage[r0 + q] = key;
      p.stage_cnt[r0 + q] = c;
      if (!p.weighted)
        { if (c < SC_SMALLHIST) atomicAdd(&s_hist[c],1u);
          else atomicAdd(p.g_hist + c,1ull);      /* c <= cap < 32767: never saturates here */
        }
      npass += (c >= p.cutoff) ? 1u : 0u;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) npass += __shfl_xor_sync(0xffffffffu,npass,o);
  if ((threadIdx.x & 31) == 0 && npass) atomicAdd(&s_pass,npass);
  __syncthreads();
  for (u32 i = threadIdx.x; i < SC_SMALLHIST; i += SC_TPB)
    { u32 c = s_hist[i];
      if (c) atomicAdd(p.g_hist + i,(u64) c);
    }
  if (threadIdx.x == 0)
    { p.e_all[g] = D;
      p.e_pass[g] = s_pass;
      if (!p.weighted) atomicAdd(p.g_ndistinct,(u64) D);
    }
}

/* ---------------------------------------------------------------------------------------------- */
/*  k_autorefine: fallback for oversize items.  One CTA per segment: finds the common prefix length of the
 *  segment's keys, then does one 8-bit MSD pass right after it, src -> dst (same offsets).  child[s*257+d]
 *  receives the child start offsets (257th = end), pcl[s] the prefix length (>= 64*NW means "all equal").  */

#define AR_TPB 512
template<int NW>
__global__ void __launch_bounds__(AR_TPB) k_autorefine(const void *buf0, void *buf1, const u64 *sstart, const u64 *send,
                                                       const u32 *sflags, u64 *child, u32 *pcl)
{ __shared__ u32 s_cnt[256];
  __shared__ u32 s_or[2*NW];
  const int s = blockIdx.x;
  const u64 start = sstart[s], n = send[s] - start;
  const bool alt = (sflags[s] & ITEM_ALTBUF) != 0;
  const Key<NW> *in = (const Key<NW> *) (alt ? buf1 : buf0) + start;
  Key<NW>      *out = (Key<NW> *) (alt ? (void *) buf0 : buf1) + start;
  if (threadIdx.x < 256) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x < 2*NW) s_or[threadIdx.x] = 0;
  __syncthreads();
  const Key<NW> k0 = in[0];
  u32 orw[2*NW];
#pragma unroll
  for (int m = 0; m < 2*NW; m++) orw[m] = 0;
  for (u64 i = threadIdx.x; i < n; i += AR_TPB)
    { Key<NW> key = in[i];
#pragma unroll
      for (int m = 0; m < NW; m++)
        { u64 x = key.w[m] ^ k0.w[m];
          orw[2*m] |= (u32) (x >> 32); orw[2*m+1] |= (u32) x;
        }
    }
#pragma unroll
  for (int m = 0; m < 2*NW; m++)
    { u32 x = __reduce_or_sync(0xffffffffu,orw[m]);
      if ((threadIdx.x & 31) == 0 && x) atomicOr(&s_or[m],x);
    }
  __syncthreads();
  int pc = 0;
  { bool done = false;
#pragma unroll
    for (int m = 0; m < 2*NW; m++)
      if (!done)
        { u32 x = s_or[m];
          if (x) { pc += __clz(x); done = true; }
          else pc += 32;
        }
  }
  if (threadIdx.x == 0) pcl[s] = (u32) pc;
  if (pc >= 64*NW)
    { if (threadIdx.x == 0) { child[(u64) s*257] = start; child[(u64) s*257+256] = start + n; }
      return;
    }
  for (u64 i = threadIdx.x; i < n; i += AR_TPB)
    atomicAdd(&s_cnt[key_digit<NW>(in[i],pc,8)],1u);
  __syncthreads();
  if (threadIdx.x == 0)
    { u32 run = 0;
      for (int d = 0; d < 256; d++)
        { u32 c = s_cnt[d];
          s_cnt[d] = run;
          child[(u64) s*257 + d] = start + run;
          run += c;
        }
      child[(u64) s*257 + 256] = start + n;
    }
  __syncthreads();
  for (u64 i = threadIdx.x; i < n; i += AR_TPB)
    { Key<NW> key = in[i];
      u32 ps = atomicAdd(&s_cnt[key_digit<NW>(key,pc,8)],1u);
      out[ps] = key;
    }
}

__global__ void k_suboff(u64 *out_off, const u32 *parent, const u64 *base, const u64 *poff, long long n)
{ long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out_off[i] = poff[parent[i]] + base[i];
}

/* ---------------------------------------------------------------------------------------------- */
/*  k_compact: staged entries of every item -> table records [kbytes key][u16 LE count], count >= cutoff.
 *  One warp per item, grid-stride.                                                                   */

struct CompactParams
  { const void *stage0; const void *stage1;
    const u32  *stage_cnt;
    const u64  *starts;
    const u32  *flags;
    const u32  *e_all;
    const u64  *out_off;        /* per item: index of its first table record */
    uint8_t    *out;
    long long   nitems;
    u32         cutoff;
    int         kbytes;
  };

template<int NW>
__global__ void __launch_bounds__(256) k_compact(CompactParams p)
{ const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
  const int tw = p.kbytes + 2;
  for (long long g = (((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5); g < p.nitems; g += nwarps)
    { const u32 D = p.e_all[g];
      if (D == 0) continue;
      const u32 fl = p.flags ? p.flags[g] : 0u;
      const Key<NW> *stage = (const Key<NW> *) ((fl & ITEM_ALTBUF) ? p.stage1 : p.stage0) + p.starts[g];
      const u32 *cnt = p.stage_cnt + p.starts[g];
      u64 o = p.out_off[g];
      for (u32 q0 = 0; q0 < D; q0 += 32)
        { u32 q = q0 + lane;
          u32 c = (q < D) ? cnt[q] : 0u;
          bool keep = (q < D) && (c >= p.cutoff);
          u32 m = __ballot_sync(0xffffffffu,keep);
          if (keep)
            { Key<NW> key = stage[q];
              uint8_t *e = p.out + (o + __popc(m & ((1u << lane) - 1u))) * (u64) tw;
              for (int b = 0; b < p.kbytes; b++)
                e[b] = (uint8_t) (key.w[b >> 3] >> (56 - 8*(b & 7)));
              e[p.kbytes]   = (uint8_t) (c & 0xffu);
              e[p.kbytes+1] = (uint8_t) (c >> 8);
            }
          o += __popc(m);
        }
    }
}

/* ============================================================================================== */
/*  Super-mer front end (the reference's own idea, split.c:1016-1393 / Appendix D of SURVEY.md, re-cut for a GPU):
 *  consecutive k-mers that share a canonical minimizer travel together as ONE 8-byte record that points into the
 *  packed reads, which stay resident in HBM:
 *      [bucket:<=24][len-1:6][strand:1][position of the first base:>=32]
 *  so the partition passes move < 1 B per k-mer instead of 16, and the counting kernel gathers the bases itself.
 *  Every instance of a canonical k-mer has the same minimizer (the minimum over BOTH strands' m-mers under a
 *  bijective order hash), hence the same bucket: each bucket is counted on chip with no cross-bucket merge.
 *  A super-mer starts at a legal k-mer that does not continue its predecessor (other bucket / illegal) or sits on
 *  a multiple of 64 (so no super-mer exceeds 64 k-mers and every thread can decide its starts locally).          */

#define SUP_PBITS_MIN 32                            /* position field width (run time, SuperGeom.pbits): global position over all ranks' read streams */
#define SUP_LMAX  64
#define SUP_BBITS 24                                /* most bucket-id bits (bucket + 6 + 1 + position bits <= 64; two partition levels of <= 11 + 13 bits) */
#define SUP_LBITS 6
#define SUP_MAXRANKS 8                              /* read streams (one per GPU of the node) a record can point into */

struct SuperParams
  { const u32 *seq; const u32 *val;
    long long  npos, nseqw, nvalw;
    int        k, m, w, p2, lmax, bbits;      /* w = k-m+1 window of m-mers, p2 = largest power of two <= w */
    int        pbits;                         /* width of the position field of a record                    */
    u64       *out; u64 cap;
    u64       *counter;                       /* [0] records emitted, [1] k-mers covered                    */
    u64        pos_offset;                    /* global position of this stream's position 0 (multi-GPU: sum of the lower ranks' lengths) */
  };

__device__ __forceinline__ u32 mix32(u32 x)   /* bijective (murmur3 finaliser) */
{ x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

#define SUP_ROWS  (SCAN_TPB + 2)                    /* rows of 32 positions: the tile + the (w-1) <= 40 position halo         */
#define SUP_RS    33                                /* row stride in words: a thread's walk along its own row is conflict-free */

/*  v[i] = min of the order keys of positions 32*row + i .. 32*row + i + P2 - 1  (valid for i < 32)  */
template<int P2>
__device__ __forceinline__ void window_min_row(const u32 *s_h, int row, u32 *v)
{ constexpr int NV = 32 + P2 - 1;
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = s_h[(row + (i >> 5))*SUP_RS + (i & 31)];
#pragma unroll
  for (int st = 1; st < P2; st <<= 1)
    {
#pragma unroll
      for (int i = 0; i < NV - st; i++) v[i] = v[i] < v[i+st] ? v[i] : v[i+st];
    }
}

/*  One CTA per tile of SCAN_TILE positions, one thread per 32 consecutive k-mer starts.  The order keys of the canonical
 *  m-mers go through shared memory once; the sliding-window minimum (width w = P2 + d, P2 <= w < 2*P2) is formed in
 *  registers by doubling up to P2 and one combine at distance d; bucket ids, run detection and the hand-over to the next
 *  thread's leading run stay in registers / warp shuffles.                                                            */
template<int P2>
__global__ void __launch_bounds__(SCAN_TPB) k_super(SuperParams p)
{ extern __shared__ u32 s_dyn[];
  u32 *s_seq = s_dyn;
  u32 *s_val = s_seq + SCAN_SEQW;
  u32 *s_h   = s_val + SCAN_VALW;           /* [SUP_ROWS][SUP_RS] order keys, later each thread's bucket ids */
  u32 *s_c   = s_h + SUP_ROWS*SUP_RS;       /* [SUP_ROWS][SUP_RS] P2-window minima (only when d > 0)          */
  __shared__ u32 s_warp[SCAN_TPB/32];
  __shared__ u64 s_base;
  constexpr int NV = 32 + P2 - 1;
  const int t = threadIdx.x;

  ScanParams sp; sp.seq = p.seq; sp.val = p.val; sp.nseqw = p.nseqw; sp.nvalw = p.nvalw;
  scan_load_tile(sp,blockIdx.x,s_seq,s_val);
  __syncthreads();

  /* order key of the canonical m-mer at every position of the tile and its halo (each computed once per tile) */
  { const int m2 = 2*p.m, sh = 32 - m2;
    const u32 mmask = (m2 == 32) ? 0xffffffffu : ((1u << m2) - 1u);
    const int nrows = SCAN_TPB + (p.w - 1 + 31) / 32;
    for (int row = t; row < nrows; row += SCAN_TPB)
      { const u32 *s = s_seq + SCAN_LHALO + 2*row;
        const u32 W0 = s[0], W1 = s[1], W2 = s[2];
        u32 *h = s_h + row*SUP_RS;
#pragma unroll
        for (int j = 0; j < 32; j++)
          { const u32 x = (j < 16) ? __funnelshift_l(W1,W0,2*j) : __funnelshift_l(W2,W1,2*(j-16));   /* 16 bases from position j */
            const u32 f = x >> sh;
            const u32 r = rc32(x) & mmask;                   /* the first m bases reverse-complemented land in the low 2m bits */
            h[j] = (mix32(f < r ? f : r) & ~1u) | (r < f ? 1u : 0u);     /* low bit: the canonical m-mer is the reverse strand here */
          }
      }
  }
  __syncthreads();

  u32 v[NV];
  window_min_row<P2>(s_h,t,v);
  const int d = p.w - P2;
  if (d > 0)
    { u32 *c = s_c + t*SUP_RS;
#pragma unroll
      for (int j = 0; j < 32; j++) c[j] = v[j];
      if (t == 0)                                            /* the halo row: its first d entries are read by the last thread */
        { u32 hv[NV];
          window_min_row<P2>(s_h,SCAN_TPB,hv);
          u32 *ch = s_c + SCAN_TPB*SUP_RS;
#pragma unroll
          for (int j = 0; j < 32; j++) ch[j] = hv[j];
        }
    }
  __syncthreads();                                           /* every read of the order keys is done; s_c is complete */
  if (d > 0)
    {
#pragma unroll
      for (int j = 0; j < 32; j++)
        { const int q = j + d;
          const u32 u = s_c[(t + (q >> 5))*SUP_RS + (q & 31)];
          v[j] = v[j] < u ? v[j] : u;
        }
    }
  /* bucket id of every k-mer start (kept in the thread's own row for the emission loop's dynamic lookups) and the strand
     on which its minimizer is canonical (the low bit of the window minimum)                                            */
  u32 *bk = s_h + t*SUP_RS;
  u32 smask = 0;                                    /* bit 31-j = minimizer of k-mer j sits on the reverse strand */
#pragma unroll
  for (int j = 0; j < 32; j++)
    { smask |= (v[j] & 1u) << (31-j);
      v[j] = ((v[j] & ~1u) * 0x9E3779B1u) >> (32 - p.bbits);
      bk[j] = v[j];
    }
  V96 vv; vv.a = s_val[t]; vv.b = s_val[t+1]; vv.c = s_val[t+2];
  const u32 ok = window_ok(vv,p.k);                 /* bit 31-j = k-mer j of this thread is legal */
  const u32 lanei = t & 31;

  /*  A super-mer is a maximal run of consecutive legal k-mers with the same bucket id, cut every SUP_LMAX k-mers counted from
      ITS OWN START and at warp boundaries (1024 positions): up to those rare cuts its extent is a property of the sequence,
      not of where the read sits in the stream, so the copies of a genomic locus in different reads give IDENTICAL super-mers
      (the bucket kernel counts duplicates once, with a weight: the reference's Supermer_Sort idea, MSDsort.c:458-489).
      cont bit j: k-mer j continues the super-mer of k-mer j-1.                                                           */
  u32 same = 0;
#pragma unroll
  for (int j = 1; j < 32; j++) same |= (v[j] == v[j-1] ? 1u : 0u) << (31-j);
  const u32 pb  = __shfl_up_sync(0xffffffffu,v[31],1);
  const u32 pok = __shfl_up_sync(0xffffffffu,ok & 1u,1);
  if (lanei != 0 && pok && pb == v[0]) same |= 0x80000000u;
  const u32 okprev = (ok >> 1) | ((lanei != 0) ? (pok << 31) : 0u);
  u32 cont = ok & okprev & same;
  { /* forced cuts of runs longer than SUP_LMAX.  rs = warp-local index of the k-mer that starts the run my k-mer 0 continues:
       it sits in the nearest lane to the left that is not wholly inside that run, as the start of its last run.          */
    const u32 lead0 = (ok >> 31) ? (u32) (1 + __clz(~(cont << 1))) : 0u;          /* my leading run: k-mers 0 .. lead0-1 */
    const bool c0 = (cont >> 31) != 0u;
    const u32 inside = __ballot_sync(0xffffffffu,c0 && lead0 == 32);
    const u32 tc = (~cont) ? (u32) (__ffs(~cont) - 1) : 32u;                      /* trailing k-mers that continue their predecessor */
    const u32 tailstart = 31u - (tc < 32u ? tc : 31u);                            /* start of the run that reaches my last k-mer */
    const u32 below = ~inside & ((1u << lanei) - 1u);
    const u32 sl = below ? (u32) (31 - __clz(below)) : 0u;
    const u32 ts = __shfl_sync(0xffffffffu,tailstart,sl);
    if (c0)
      { const u32 d0 = 32u*lanei - (32u*sl + ts);                                 /* k-mers of the run before my window: >= 1 */
        const u32 off = (SUP_LMAX - (d0 & (SUP_LMAX-1))) & (SUP_LMAX-1);          /* first j with (d0 + j) a multiple of SUP_LMAX */
        if (off < lead0) cont &= ~(0x80000000u >> off);
      }
  }
  const u32 starts = ok & ~cont;                    /* bit 31-j = a super-mer starts at j */
  /* how far a super-mer that reaches the end of my window runs on: leading runs of the next two lanes (<= SUP_LMAX in all) */
  const u32 lead = (ok >> 31) ? (u32) (1 + __clz(~(cont << 1))) : 0u;
  const u32 c0f = cont >> 31;
  u32 n1lead = __shfl_down_sync(0xffffffffu,lead,1), n1c0 = __shfl_down_sync(0xffffffffu,c0f,1);
  u32 n2lead = __shfl_down_sync(0xffffffffu,lead,2), n2c0 = __shfl_down_sync(0xffffffffu,c0f,2);
  if (lanei >= 31) n1c0 = 0;
  if (lanei >= 30) n2c0 = 0;
  const u32 ext1 = n1c0 ? n1lead : 0u;
  const u32 ext  = ext1 + ((ext1 == 32u && n2c0) ? n2lead : 0u);

  const u32 nrun = __popc(starts);
  u32 incl = nrun;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
    { u32 y = __shfl_up_sync(0xffffffffu,incl,o);
      if ((t & 31) >= o) incl += y;
    }
  if ((t & 31) == 31) s_warp[t >> 5] = incl;
  u32 nk = __popc(ok);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nk += __shfl_xor_sync(0xffffffffu,nk,o);
  __syncthreads();
  u32 woff = 0, tot = 0;
  for (int i = 0; i < SCAN_TPB/32; i++)
    { u32 x = s_warp[i];
      if (i < (t >> 5)) woff += x;
      tot += x;
    }
  if ((t & 31) == 0 && nk) atomicAdd(p.counter + 1,(u64) nk);
  if (t == 0) s_base = tot ? atomicAdd(p.counter,(u64) tot) : 0ull;
  __syncthreads();
  u64 pos = s_base + woff + incl - nrun;
  if (s_base + tot > p.cap) return;                      /* buffer too small: the host sees counter > cap and falls back */

  const u64 tile0 = p.pos_offset + (u64) blockIdx.x * SCAN_TILE;
  const int base = t * SCAN_PPT;
  u32 todo = starts;
  while (todo)
    { const int j = __clz(todo);
      todo &= ~(0x80000000u >> j);
      const u32 after = (j == 31) ? 0u : (0xffffffffu >> (j+1));
      const u32 stop = ~cont & after;
      const int e = stop ? __clz(stop) : SCAN_PPT;       /* window-local end (exclusive) */
      u32 len = (u32) (e - j);
      if (e == SCAN_PPT) len += ext;                     /* runs on into the next lanes' windows */
      if (len > SUP_LMAX) len = SUP_LMAX;                /* cannot happen (forced cuts); keeps the length field sound */
      const u32 b = bk[j];
      const u32 sb = (smask >> (31-j)) & 1u;
      p.out[pos++] = ((u64) b << (64 - p.bbits)) | ((u64) (len-1) << (p.pbits+1)) | ((u64) sb << p.pbits) | (tile0 + (u64) (base + j));
    }
}

/*  fields of a super-mer record [bucket : bbits][# k-mers - 1 : 6][strand : 1][position of the first base : pbits]  */
__device__ __forceinline__ u32 sm_len(u64 sm, int pbits)    { return (u32) ((sm >> (pbits+1)) & 63ull) + 1u; }
__device__ __forceinline__ u32 sm_strand(u64 sm, int pbits) { return (u32) (sm >> pbits) & 1u; }

struct BucketParams
  { const u64 *recs; const u32 *seq;
    const u32 *seqr[SUP_MAXRANKS];                       /* multi-GPU: packed reads of every rank (peer memory over NVLink) */
    u64        pbase[SUP_MAXRANKS];                      /* global position of rank r's position 0                          */
    int        nranks;                                   /* 1: every record points into seq                                 */
    int        pbits;                                    /* width of the position field of a record                         */
    const uint4 *payload;                                /* PAY instantiation: position field = index of the super-mer's 32-byte left-aligned
                                                            base string in this array (multi-GPU: exchanged with the records); the
                                                            position field is a WORD offset into it                               */
    const u64 *starts; const u64 *ends; long long nitems;
    int        k;
    u64       *g_hist; u64 *g_maxinst; u64 *g_ndistinct;
    void      *ent; u64 ent_cap; u64 *ent_counter;       /* distinct entries out (may be NULL): Key<2> (key | count), or Key<3> (key, count) when k > 56 */
    u32        ent_min;                                  /* only entries with (saturated) count >= ent_min are emitted */
    u32       *g_fail;                                   /* [0] set if a group could not be counted, [1] # of hash classes split */
    /* groups of more than `big` super-mers (a giant bucket: a high-copy repeat, a low-complexity run) are not counted on chip:
       one CTA / warp would grind through them alone.  They are listed here and counted by the record pipeline instead.     */
    u32        big;
    u32       *spill_cnt; u32 *spill_list; u32 spill_cap;
    u64       *spill_kmers;                              /* k-mers covered by the listed groups */
    u64       *g_stat;                                   /* [0] super-mers met by the bucket kernel, [1] of them expanded (the others were copies) */
  };

/*  Strand arithmetic on KW = ceil(2k/32) 32-bit words (most significant first; the 2k key bits left aligned, the
 *  s = 32*KW - 2k < 32 low bits of the last word zero).  klast = mask of the used bits of the last word.          */

/* both strands of k-mer number j of a super-mer whose base string is b[0..] (left aligned): F = forward, G = reverse complement */
template<int KW>
__device__ __forceinline__ void supermer_strands(const u32 *b, int j, int k, u32 klast, u32 *F, u32 *G)
{ const int q = j >> 4, sh = 2*(j & 15);
#pragma unroll
  for (int t = 0; t < KW; t++) F[t] = __funnelshift_l(b[q+t+1],b[q+t],sh);
  F[KW-1] &= klast;
  /* reverse complement of the 16*KW base slots, then drop the 16*KW - k padding slots off the top */
  u32 Z[KW+1];
#pragma unroll
  for (int t = 0; t < KW; t++) Z[t] = rc32(F[KW-1-t]);
  Z[KW] = 0;
  const int s = 32*KW - 2*k;
#pragma unroll
  for (int t = 0; t < KW; t++) G[t] = __funnelshift_l(Z[t+1],Z[t],s);
  G[KW-1] &= klast;
}

/* slide both strands one base to the right: base code c enters the forward strand at its low end */
template<int KW>
__device__ __forceinline__ void strands_roll(u32 *F, u32 *G, u32 c, int s /* 32*KW - 2k */, u32 klast)
{
#pragma unroll
  for (int t = 0; t < KW-1; t++) F[t] = __funnelshift_l(F[t+1],F[t],2);
  F[KW-1] = (F[KW-1] << 2) | (c << s);
#pragma unroll
  for (int t = KW-1; t > 0; t--) G[t] = __funnelshift_r(G[t],G[t-1],2);
  G[0] = (G[0] >> 2) | ((3u - c) << 30);
  G[KW-1] &= klast;
}

/* canonical key = the smaller strand, as the 16-byte record key (words beyond KW are zero) */
template<int KW>
__device__ __forceinline__ Key<2> strands_canon(const u32 *F, const u32 *G)
{ const u64 f0 = ((u64) F[0] << 32) | F[1], g0 = ((u64) G[0] << 32) | G[1];
  u64 f1 = 0, g1 = 0;
  if (KW > 2) { f1 = (u64) F[2] << 32; g1 = (u64) G[2] << 32; }
  if (KW > 3) { f1 |= F[KW > 3 ? 3 : 0]; g1 |= G[KW > 3 ? 3 : 0]; }
  const bool lt = (g0 < f0) | ((g0 == f0) & (g1 < f1));
  Key<2> key;
  key.w[0] = lt ? g0 : f0;
  key.w[1] = lt ? g1 : f1;
  return key;
}

/* slot / round hash of a canonical key: one multiply-add per word and a finaliser */
template<int KW>
__device__ __forceinline__ u32 bucket_hash(const Key<2> &a)
{ u32 h = (u32) (a.w[0] >> 32) * 0x9E3779B1u + (u32) a.w[0] * 0x85EBCA77u;
  if (KW > 2) h += (u32) (a.w[1] >> 32) * 0xC2B2AE3Du;
  if (KW > 3) h += (u32) a.w[1] * 0x27D4EB2Fu;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 16;
  return h;
}

/*  multi-GPU exchange: the 32-byte left-aligned base string (<= 64 + k - 1 <= 128 bases) of every super-mer record, in record
 *  order, gathered out of the sender's own packed reads -- it travels beside the 8-byte records in the all-to-all.       */
__global__ void __launch_bounds__(256) k_materialise(const u64 *recs, long long n, int pbits, u64 pos_offset, int k, const u32 *seq, uint4 *payload)
{ const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 sm = recs[i];
  const u32 l = sm_len(sm,pbits);
  const u64 ps = (sm & ((1ull << pbits) - 1ull)) - pos_offset;
  const u32 *g = seq + (ps >> 4);
  const int sh = 2*(int) (ps & 15ull);
  const int nw = (int) ((2*(l + k - 1) + sh + 31) >> 5);
  u32 x[9], d[8];
#pragma unroll
  for (int t = 0; t < 9; t++) x[t] = (t < nw) ? __ldg(g + t) : 0u;
#pragma unroll
  for (int t = 0; t < 8; t++) d[t] = __funnelshift_l(x[t+1],x[t],sh);
  payload[2*i]   = make_uint4(d[0],d[1],d[2],d[3]);
  payload[2*i+1] = make_uint4(d[4],d[5],d[6],d[7]);
}

/*  after the exchange the position field of received record i becomes the WORD offset of its base string in the payload:
    8*i for the fixed 32-byte strings of k_materialise                                                                     */
__global__ void __launch_bounds__(256) k_reindex(u64 *recs, long long n, int pbits)
{ const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) recs[i] = (recs[i] & ~((1ull << pbits) - 1ull)) | (u64) (8*i);
}

/*  Compact payload of the in-library exchange: a super-mer's base string travels as ceil((l + k - 1) / 16) words (15 bytes
    on average at k = 40 instead of 32).  k_payload_words: words of every record; after an exclusive scan (woff),
    k_materialise_compact writes the strings back to back and the record to send, whose position field becomes the word
    offset of its string INSIDE ITS DESTINATION'S SLICE (slice r = records [sstart[r], sstart[r+1]) ); the receiver adds
    the word offset at which the slice of every source landed (k_rebase_slices).                                          */
__global__ void __launch_bounds__(256) k_payload_words(const u64 *recs, long long n, int pbits, int k, u32 *nw)
{ const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) nw[i] = (sm_len(recs[i],pbits) + (u32) k - 1u + 15u) >> 4;
}

struct SliceTable { u64 start[SUP_MAXRANKS*2 + 1]; u64 add[SUP_MAXRANKS*2 + 1]; int n; };

__global__ void __launch_bounds__(256) k_materialise_compact(const u64 *recs, long long n, int pbits, u64 pos_offset, int k, const u32 *seq,
                                                             const u64 *woff, SliceTable st, u64 *send, u32 *payload)
{ const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 sm = recs[i];
  const u32 l = sm_len(sm,pbits);
  const u64 pmask = (1ull << pbits) - 1ull;
  const u64 ps = (sm & pmask) - pos_offset;
  const u32 *g = seq + (ps >> 4);
  const int sh = 2*(int) (ps & 15ull);
  const int nwr = (int) ((2*(l + k - 1) + sh + 31) >> 5);       /* packed words the string touches in the reads */
  const int nwo = (int) ((l + (u32) k - 1u + 15u) >> 4);         /* words it occupies left aligned               */
  u32 x[9];
#pragma unroll
  for (int t = 0; t < 9; t++) x[t] = (t < nwr) ? __ldg(g + t) : 0u;
  const u64 w0 = woff[i];
#pragma unroll
  for (int t = 0; t < 8; t++)
    if (t < nwo) payload[w0 + t] = __funnelshift_l(x[t+1],x[t],sh);
  int r = 0;
#pragma unroll 1
  for (int q = 1; q < st.n; q++)
    if ((u64) i >= st.start[q]) r = q;
  send[i] = (sm & ~pmask) | (w0 - woff[st.start[r]]);
}

/*  received records [start[r], start[r+1]) came from source r, whose strings landed at word offset add[r]  */
__global__ void __launch_bounds__(256) k_rebase_slices(u64 *recs, long long n, SliceTable st)
{ const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = 0;
#pragma unroll 1
  for (int q = 1; q < st.n; q++)
    if ((u64) i >= st.start[q]) r = q;
  recs[i] += st.add[r];
}

/* k-mers covered by the records of every level-1 bucket b = [off[b], off[b+1]): one CTA per bucket (plans the bucket-range rounds) */
__global__ void __launch_bounds__(256) k_bucket_kmers(const u64 *recs, const u64 *off, int pbits, u64 *out)
{ __shared__ u64 s_w[8];
  const u64 a = off[blockIdx.x], b = off[blockIdx.x + 1];
  u64 s = 0;
  for (u64 i = a + threadIdx.x; i < b; i += 256) s += sm_len(recs[i],pbits);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu,s,o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0)
    { u64 t = 0;
      for (int i = 0; i < 8; i++) t += s_w[i];
      out[blockIdx.x] = t;
    }
}

/* total k-mers covered by n super-mer records (sizes the distinct-entry buffer of a rank after the exchange) */
__global__ void __launch_bounds__(256) k_sum_lengths(const u64 *recs, long long n, int pbits, u64 *total)
{ u64 s = 0;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
    s += sm_len(recs[i],pbits);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu,s,o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(total,s);
}

/* ---------------------------------------------------------------------------------------------- */
/*  Count profiles (-p).  The finished distinct-k-mer table (all counts, no cutoff) is compacted into
 *  (keys[], cnts[]) with a 2^B-slot prefix index; k_profile re-scans the packed reads and looks every
 *  canonical k-mer up -- the reference's "relative profile" sort-merge-join idea (count.c:675-792)
 *  turned into an indexed lookup.  Result: one u16 per read position (0 where no legal k-mer starts).  */

template<int NW>
__global__ void __launch_bounds__(256) k_compact_keys(CompactParams p, Key<(NW == 3) ? 2 : NW> *keys, uint16_t *cnts)
{ const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
  for (long long g = (((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5); g < p.nitems; g += nwarps)
    { const u32 D = p.e_all[g];
      if (D == 0) continue;
      const u32 fl = p.flags ? p.flags[g] : 0u;
      const Key<NW> *stage = (const Key<NW> *) ((fl & ITEM_ALTBUF) ? p.stage1 : p.stage0) + p.starts[g];
      const u32 *cnt = p.stage_cnt + p.starts[g];
      const u64 o = p.out_off[g];
      for (u32 q = lane; q < D; q += 32)
        { const Key<NW> sk = stage[q];
          Key<(NW == 3) ? 2 : NW> ok;
#pragma unroll
          for (int m = 0; m < ((NW == 3) ? 2 : NW); m++) ok.w[m] = sk.w[m];
          keys[o+q] = ok;
          cnts[o+q] = (uint16_t) cnt[q];
        }
    }
}

/*  Merging tables (Fastmerge.c:168-450 on the device): the records of all tables, sorted together, hold runs of equal keys.
 *  k_run_heads flags the first record of every run; k_merge_runs (one thread per record, heads only) adds the counts of its
 *  run -- saturating at 32767, histogram of the merged counts, and for a saturated sum the instances its unsaturated
 *  members stood for go to max_inst (Fastmerge.c:311-331) -- and writes the merged record at its rank among the heads.     */
__global__ void __launch_bounds__(256) k_run_heads(const uint8_t *tab, u64 n, int kbytes, u32 *head)
{ const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int tw = kbytes + 2;
  bool h = (i == 0);
  if (!h)
    { const uint8_t *a = tab + i * (u64) tw, *b = a - tw;
      for (int x = 0; x < kbytes; x++) h = h || (a[x] != b[x]);
    }
  head[i] = h ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_merge_runs(const uint8_t *tab, u64 n, int kbytes, const u32 *head, const u64 *pos, uint8_t *out,
                                                    u64 *g_hist, u64 *g_maxinst)
{ const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !head[i]) return;
  const int tw = kbytes + 2;
  const uint8_t *e = tab + i * (u64) tw;
  u64 sum = 0, small = 0;
  for (u64 j = i; j < n && (j == i || !head[j]); j++)
    { const uint8_t *r = tab + j * (u64) tw;
      const u32 c = (u32) r[kbytes] | ((u32) r[kbytes+1] << 8);
      sum += c;
      if (c < 0x7fffu) small += c;
    }
  const u32 sc = sum > 0x7fffull ? 0x7fffu : (u32) sum;
  atomicAdd(g_hist + sc,1ull);
  if (sum > 0x7fffull && small) atomicAdd(g_maxinst,small);
  uint8_t *o = out + pos[i] * (u64) tw;
  for (int x = 0; x < kbytes; x++) o[x] = e[x];
  o[kbytes] = (uint8_t) (sc & 0xffu); o[kbytes+1] = (uint8_t) (sc >> 8);
}

/*  table records [kbytes key][u16 LE count] (a .ktab as the host read it back) -> the lookup arrays of k_profile  */
template<int NW>
__global__ void __launch_bounds__(256) k_records_to_keys(const uint8_t *tab, u64 n, int kbytes, Key<NW> *keys, uint16_t *cnts)
{ const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t *e = tab + i * (u64) (kbytes + 2);
  Key<NW> r;
#pragma unroll
  for (int m = 0; m < NW; m++) r.w[m] = 0;
  for (int b = 0; b < kbytes; b++) r.w[b >> 3] |= (u64) e[b] << (56 - 8*(b & 7));
  keys[i] = r;
  cnts[i] = (uint16_t) ((u32) e[kbytes] | ((u32) e[kbytes+1] << 8));
}

template<int NW>
__global__ void k_build_index(const Key<NW> *keys, u64 n, int B, u64 *idx)
{ u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const u64 slots = 1ull << B;
  u64 lo = (i == 0) ? 0 : (keys[i-1].w[0] >> (64-B)) + 1;
  u64 hi = (i == n) ? slots : (keys[i].w[0] >> (64-B));
  for (u64 s = lo; s <= hi && s <= slots; s++) idx[s] = i;
}

/*  The lookup table of the profiles as a hash table: buckets of four 16-byte slots = one 64-byte line = two 32-byte sectors.
 *  A key's home is one HALF (sector) of its bucket; it is stored in the first free slot of: home half, other half, next
 *  bucket ...  A lookup reads the home half with ONE 256-bit load (LDG.E.256, new with sm_100) and in most cases is done:
 *  one DRAM sector per lookup.  (History, ncu r2/r2n: sorted keys + prefix index + counts = 3 dependent random accesses,
 *  294 B of DRAM per lookup; four 128-bit loads of a 64-byte bucket = 4 DRAM sectors, 128 B -- two loads that miss on the
 *  same sector are both sent to DRAM.)
 *  Slot = { key word 0, key word 1 | 0x8000 | count } when the key leaves the low 16 bits of word 1 free (k <= 56; for
 *  k <= 32 word 1 is the count alone), empty = second word zero.  "wide" (k 57..64): the slot holds the two key words, the
 *  count (| 0x8000) sits in hcnt[slot], empty = hcnt zero.  Keys are distinct, so building is claim-a-slot (one atomicCAS on
 *  the word that holds the count), then a plain store of the rest; lookups run after the build.                          */
struct ProfHash { const ulonglong2 *slots; const uint16_t *hcnt; u64 nbuckets; int wide; };   /* wide: bit 0 = counts in hcnt, bit 1 = lookups carry the L2::64B hint */

__device__ __forceinline__ u64 fk_mix64(u64 x)
{ x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

template<int NW> __device__ __forceinline__ u64 fk_key_hash(const Key<NW> &k)
{ u64 h = fk_mix64(k.w[0]);
  if (NW > 1) h = fk_mix64(h ^ k.w[1]);
  return h;
}

template<int NW>
__global__ void __launch_bounds__(256) k_hash_build(const Key<NW> *keys, const uint16_t *cnts, u64 n, ulonglong2 *slots, uint16_t *hcnt,
                                                    u64 nbuckets, int wide)
{ const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Key<NW> key = keys[i];
  const u64 c  = 0x8000ull | (u64) (cnts[i] & 0x7fffu);
  const u64 k1 = (NW > 1) ? key.w[1] : 0ull;
  const u64 h  = fk_key_hash<NW>(key);
  u64 b = __umul64hi(h,nbuckets);
  const u32 s0 = ((u32) h & 1u) << 1;                 /* home half: slots s0, s0+1; then s0^2, (s0^2)+1 */
  for (;;)
    { for (u32 t = 0; t < 4; t++)
        { const u64 x = 4*b + (s0 ^ t);
          if (wide)
            { if (atomicCAS((unsigned short *) hcnt + x,(unsigned short) 0,(unsigned short) c) == 0)
                { slots[x] = make_ulonglong2(key.w[0],k1); return; }
            }
          else if (atomicCAS((unsigned long long *) &slots[x].y,0ull,(unsigned long long) (k1 | c)) == 0ull)
            { slots[x].x = key.w[0]; return; }
        }
      b = (b + 1 == nbuckets) ? 0 : b + 1;
    }
}

/*  count of the key (k0,k1), 0 if absent (relative profiles look up k-mers the table never saw).  NOT inlined: k_profile
 *  unrolls its 32 positions, and 32 copies of this loop made a 164 KB kernel (ncu r2n: no_instruction 31 warps per issue). */
template<int NW>
__device__ __noinline__ u32 hash_lookup(const ulonglong2 *slots, const uint16_t *hcnt, u64 nbuckets, int wide, u64 k0, u64 k1)
{ if (NW == 1) k1 = 0ull;
  Key<NW> key; key.w[0] = k0; if (NW > 1) key.w[NW-1] = k1;
  const u64 h = fk_key_hash<NW>(key);
  u64 b = __umul64hi(h,nbuckets);
  const u32 s0 = ((u32) h & 1u) << 1;
  for (;;)
    {
#pragma unroll
      for (u32 half = 0; half < 2; half++)
        { const u64 x = 4*b + (s0 ^ (half << 1));       /* slots x, x+1: one 32-byte sector */
          u64 ax, ay, bx, by;
          if (wide & 2) asm volatile("ld.global.nc.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(ax), "=l"(ay), "=l"(bx), "=l"(by) : "l"(slots + x));
          else          asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(ax), "=l"(ay), "=l"(bx), "=l"(by) : "l"(slots + x));
          if (!(wide & 1))
            { if (ax == k0 && (ay & ~0xffffull) == k1 && (ay & 0x8000ull)) return (u32) (ay & 0x7fffull);
              if (ay == 0ull) return 0u;
              if (bx == k0 && (by & ~0xffffull) == k1 && (by & 0x8000ull)) return (u32) (by & 0x7fffull);
              if (by == 0ull) return 0u;
            }
          else
            { if (ax == k0 && ay == k1) { const u32 c = hcnt[x]; return c & 0x7fffu; }       /* c == 0: the all-zero key met a free slot */
              if ((ax | ay) == 0ull && hcnt[x] == 0) return 0u;
              if (bx == k0 && by == k1) { const u32 c = hcnt[x+1]; return c & 0x7fffu; }
              if ((bx | by) == 0ull && hcnt[x+1] == 0) return 0u;
            }
        }
      b = (b + 1 == nbuckets) ? 0 : b + 1;
    }
}

/*  k_profile<NW,HASH,DIRECT>.  DIRECT: the grid walks VIRTUAL tiles -- the tiles of the chunks of the read stream taken in
 *  the order of the output (input thread by input thread, each thread's chunks as they arrived), tile t starting at position
 *  vt_src[t] -- and every thread writes the counts of its 32 positions straight to their places in the output: pieces
 *  [vt_plo[t], vt_pend[t]) are the reads (src, len -> dst) that meet the tile, sorted by src.  Output order = launch order, so
 *  the host copies the output back in slices while later tiles are still being looked up.  !DIRECT: tile b = positions
 *  [b*SCAN_TILE ..), one u16 per position into raw[] (k_gather_profile then moves the pieces).                            */
struct ProfileParams
  { ScanParams   sp;
    const void  *keys; const uint16_t *cnts; const u64 *idx; int B;      /* !HASH: sorted keys, counts, 2^B-slot prefix index */
    ProfHash     H;
    uint16_t    *raw;            /* !DIRECT: [npos] */
    const long long *vt_src; const int *vt_plo, *vt_pend; long long vt0;  /* DIRECT: virtual tiles, first tile of this launch */
    const long long *psrc, *pdst; const int *plen;
    uint16_t    *out;
  };

__device__ __forceinline__ void scan_load_at(const ScanParams &p, long long pos0, u32 *s_seq, u32 *s_val)   /* pos0 % 32 == 0 */
{ const long long w0 = (pos0 >> 4) - SCAN_LHALO;
  for (int i = threadIdx.x; i < SCAN_SEQW; i += SCAN_TPB)
    { long long g = w0 + i;
      s_seq[i] = (g >= 0 && g < p.nseqw) ? __ldg(p.seq+g) : 0u;
    }
  const long long v0 = pos0 >> 5;
  for (int i = threadIdx.x; i < SCAN_VALW; i += SCAN_TPB)
    { long long g = v0 + i;
      s_val[i] = (g < p.nvalw) ? __ldg(p.val+g) : 0u;
    }
}

template<int NW, bool HASH, bool DIRECT>
__global__ void __launch_bounds__(SCAN_TPB) k_profile(ProfileParams q)
{ extern __shared__ u32 s_dyn[];
  u32 *s_seq = s_dyn;
  u32 *s_val = s_seq + SCAN_SEQW;
  constexpr int NW32 = 2*NW;
  const ScanParams &p = q.sp;
  const long long vt = DIRECT ? q.vt0 + blockIdx.x : (long long) blockIdx.x;
  const long long tile0 = DIRECT ? q.vt_src[vt] : vt * SCAN_TILE;
  scan_load_at(p,tile0,s_seq,s_val);
  __syncthreads();
  Window w;
  load_window(w,s_seq,s_val,threadIdx.x,p.k);
  u32 km[4] = { p.kmask[0], p.kmask[1], p.kmask[2], p.kmask[3] };
  const Key<NW> *keys = (const Key<NW> *) q.keys;
  const long long p0 = tile0 + (long long) threadIdx.x * SCAN_PPT;
  u32 res[SCAN_PPT/2];
#pragma unroll
  for (int i = 0; i < SCAN_PPT/2; i++) res[i] = 0;
  auto fn = [&](int j, const u32 *C)
    { Key<NW> key;
#pragma unroll
      for (int m = 0; m < NW; m++) key.w[m] = ((u64) C[2*m] << 32) | C[2*m+1];
      u32 c = 0;
      if (HASH) c = hash_lookup<NW>(q.H.slots,q.H.hcnt,q.H.nbuckets,q.H.wide,key.w[0],(NW > 1) ? key.w[NW-1] : 0ull);
      else
        { const u64 s = key.w[0] >> (64 - q.B);
          u64 lo = q.idx[s], hi = q.idx[s+1];
          while (lo < hi)
            { u64 mid = (lo + hi) >> 1;
              Key<NW> t = keys[mid];
              if (key_eq<NW>(t,key)) { c = q.cnts[mid]; break; }
              if (key_lt<NW>(t,key)) lo = mid+1; else hi = mid;
            }
        }
      res[j>>1] |= c << (16*(j&1));
    };
  if (!DIRECT)
    { KmerLoop<NW32,0>::run(w,km,fn);
      if (p0 < p.npos)
        { u32 *out = (u32 *) (q.raw + p0);            /* p0 is a multiple of 32: 4-byte aligned */
#pragma unroll
          for (int i = 0; i < SCAN_PPT/2; i++)
            if (p0 + 2*i < p.npos) out[i] = res[i];
        }
      return;
    }
  /* the first piece that ends beyond p0 */
  int i = q.vt_plo[vt];
  const int pe = q.vt_pend[vt];
  { int lo = i, hi = pe;
    while (lo < hi)
      { const int mid = (lo + hi) >> 1;
        if (q.psrc[mid] + q.plen[mid] > p0) hi = mid; else lo = mid + 1;
      }
    i = lo;
  }
  if (i >= pe || q.psrc[i] >= p0 + SCAN_PPT) return;          /* no read position among these 32: nothing to look up */
  KmerLoop<NW32,0>::run(w,km,fn);
  long long s = q.psrc[i], e = s + q.plen[i], db = q.pdst[i] - s;   /* output index of position x of piece i: db + x */
  if (p0 >= s && p0 + SCAN_PPT <= e)
    { uint16_t *d = q.out + (db + p0);
      if ((((size_t) d) & 3) == 0)
        { u32 *d32 = (u32 *) d;
#pragma unroll
          for (int x = 0; x < SCAN_PPT/2; x++) d32[x] = res[x];
        }
      else                                               /* odd output index: the pairs straddle the words */
        { d[0] = (uint16_t) (res[0] & 0xffffu);
          u32 *d32 = (u32 *) (d + 1);
#pragma unroll
          for (int x = 0; x + 1 < SCAN_PPT/2; x++) d32[x] = __funnelshift_r(res[x],res[x+1],16);
          d[SCAN_PPT-1] = (uint16_t) (res[SCAN_PPT/2-1] >> 16);
        }
      return;
    }
#pragma unroll
  for (int j = 0; j < SCAN_PPT; j++)
    { const long long pos = p0 + j;
      while (i < pe && pos >= e)
        { i += 1;
          if (i < pe) { s = q.psrc[i]; e = s + q.plen[i]; db = q.pdst[i] - s; }
        }
      if (i < pe && pos >= s) q.out[db + pos] = (uint16_t) ((res[j>>1] >> (16*(j&1))) & 0xffffu);
    }
}

/*  piece r: copy raw[src[r] .. src[r]+len[r]) to out[dst[r] ..]; one warp per piece                 */
__global__ void k_gather_profile(const uint16_t *raw, const long long *src, const long long *dst, const int *len,
                                 long long npieces, uint16_t *out)
{ const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
  for (long long r = (((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < npieces; r += nwarps)
    { const uint16_t *s = raw + src[r];
      uint16_t *d = out + dst[r];
      const int n = len[r];
      for (int i = lane; i < n; i += 32) d[i] = s[i];
    }
}

}  // namespace fk
