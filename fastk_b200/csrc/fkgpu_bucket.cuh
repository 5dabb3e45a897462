/*  fkgpu_bucket.cuh -- k_bucket_count: on-chip expansion + hash count of whole minimizer buckets (sm_100a).
 *
 *  Replaces, for the k-mers of one bucket group, the reference's  supermer_list_thread -> Supermer_Sort -> kmer_list_thread
 *  -> Weighted_Kmer_Sort / hist_kmers  chain (count.c:165-313, MSDsort.c:458-489, count.c:339-542, MSDsort.c:491-544):
 *  every instance of a canonical k-mer lives in one bucket, so a bucket group is counted completely inside one CTA.
 *
 *  Persistent CTAs walk the work groups (whole buckets, ~BK_T super-mers).  Per group and hash class:
 *    load     every warp owns an equal slice of the group's super-mer records (<= 32 per piece, one per lane): the lane
 *             gathers its super-mer's bases from the resident packed reads (or a peer's HBM / the exchanged payload) and
 *             parks them left-aligned in the warp's own rows of shared memory -- no block barrier, only __syncwarp
 *    expand   the warp's k-mer instances are dealt out 32 per round, ONE PER LANE: the lane finds its (super-mer, offset)
 *             from a warp prefix sum of the lengths with one REDUX.OR + popc + shuffle, cuts both strands out of the row
 *             (funnel shifts) and takes the smaller -- no per-thread binary search, no rolling state, no divergence at
 *             super-mer boundaries
 *    count    open-addressing table in shared memory: slot word = [16-bit fingerprint | key index + 1]; the distinct
 *             keys live in SoA arrays k0[] / k1[] with their counts in cnt[]; a probe compares the fingerprint before it
 *             touches the key, a first occurrence claims the slot with one CAS
 *    emit     histogram contributions (CTA-private small histogram, flushed once per CTA), max_inst, and the distinct
 *             (key | saturated count) entries with count >= the table cutoff, appended at ONE global atomic per class
 *  A class whose distinct keys overflow the pool (or whose probes run long) is split in two residue classes of the key
 *  hash and redone (binary tree walked without a stack).                                                             */
#pragma once
#include "fkgpu_kernels.cuh"

namespace fk {

#define BK_TPB    256
#define BK_WARPS  (BK_TPB/32)
#define BK_GC     (BK_WARPS*32)      /* super-mers per piece: one per lane                                   */
#define BK_DC     1024               /* distinct keys a class may hold                                       */
#define BK_TS     2048               /* slots (load <= 0.5)                                                   */
#define BK_ROW    8                  /* 32-bit words of a super-mer's base string (<= 64 + k - 1 <= 127 bases) */
#define BK_PROBE  96                 /* probe steps after which a class is split rather than ground through   */
#define BK_MAXR   4096               /* most residue classes a group is split into                            */

#define BK_SMEM   ((size_t) BK_DC*8*2 + (size_t) BK_DC*4 + (size_t) BK_TS*4 + (size_t) BK_GC*BK_ROW*4)

/*  Super-mer record -> its base string (l + k - 1 bases, 2 bits each) left aligned in d[0..8).  With `orient` the bits beyond
 *  the string are cleared and the string is reverse-complemented when the record says its minimizer is canonical on the
 *  reverse strand: copies of one genomic locus, read from either strand, then give the SAME string (and the same length), which
 *  is what lets the bucket kernel count duplicate super-mers once.                                                        */
template<bool PAY>
__device__ __forceinline__ void load_supermer(const BucketParams &p, u64 sm, u64 pmask, u32 l, bool orient, u32 *d)
{ u64 ps = sm & pmask;
  if (PAY)
    { /* exchanged base strings: left aligned, ps = word offset, only the words the string occupies are there */
      const u32 *pw = (const u32 *) p.payload + ps;
      const int nwo = (int) ((l + (u32) p.k - 1u + 15u) >> 4);
#pragma unroll
      for (int t = 0; t < 8; t++) d[t] = (t < nwo) ? __ldg(pw + t) : 0u;
    }
  else
    { const u32 *sq = p.seq;
      if (p.nranks > 1)
        { int r = 0;                    /* owner = last rank whose base is <= ps; its stream may live on a peer GPU */
#pragma unroll 1
          for (int q = 1; q < p.nranks; q++)
            if (ps >= p.pbase[q]) r = q;
          ps -= p.pbase[r]; sq = p.seqr[r];
        }
      const u32 *gp = sq + (ps >> 4);
      const int sh = 2*(int) (ps & 15ull);
      const int nw = (int) ((2*(l + p.k - 1) + sh + 31) >> 5);          /* packed words this super-mer touches */
      u32 x[9];
#pragma unroll
      for (int t = 0; t < 9; t++) x[t] = (t < nw) ? __ldg(gp + t) : 0u;
#pragma unroll
      for (int t = 0; t < 8; t++) d[t] = __funnelshift_l(x[t+1],x[t],sh);
    }
  if (orient)
    { const u32 nb = l + (u32) p.k - 1u;
      const u32 wq = nb >> 4, wb = 2u*(nb & 15u);
#pragma unroll
      for (int t = 0; t < 8; t++)
        d[t] = ((u32) t < wq) ? d[t] : (((u32) t == wq && wb) ? (d[t] & (0xffffffffu << (32u - wb))) : 0u);
      if (sm_strand(sm,p.pbits))
        { /* reverse complement of the 128-base slot, then drop the 128 - nb pad bases that lead it: a left shift by a
             run-time number of words (three select stages, all in registers) and bits (funnel shifts)                */
          u32 z[8];
#pragma unroll
          for (int t = 0; t < 8; t++) z[t] = rc32(d[7-t]);
          const u32 sh2 = 2u*(128u - nb), ws = sh2 >> 5, bs = sh2 & 31u;
#pragma unroll
          for (int t = 0; t < 8; t++) z[t] = (ws & 4u) ? ((t + 4 < 8) ? z[(t + 4) & 7] : 0u) : z[t];
#pragma unroll
          for (int t = 0; t < 8; t++) z[t] = (ws & 2u) ? ((t + 2 < 8) ? z[(t + 2) & 7] : 0u) : z[t];
#pragma unroll
          for (int t = 0; t < 8; t++) z[t] = (ws & 1u) ? ((t + 1 < 8) ? z[(t + 1) & 7] : 0u) : z[t];
#pragma unroll
          for (int t = 0; t < 8; t++) d[t] = __funnelshift_l((t + 1 < 8) ? z[(t + 1) & 7] : 0u,z[t],bs);
        }
    }
}

/*  one warp lists an oversize group for the record pipeline and adds up the k-mers its super-mers cover */
__device__ __forceinline__ void spill_group(const BucketParams &p, long long g, u64 r0, u64 r1, u32 lane)
{ u64 s = 0;
  for (u64 i = r0 + lane; i < r1; i += 32) s += sm_len(p.recs[i],p.pbits);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu,s,o);
  if (lane == 0)
    { const u32 at = atomicAdd(p.spill_cnt,1u);
      if (at < p.spill_cap) p.spill_list[at] = (u32) g;
      atomicAdd(p.spill_kmers,s);
    }
}

template<int KW, bool PAY, bool WIDE>
__global__ void __launch_bounds__(BK_TPB,4) k_bucket_count2(BucketParams p, u32 klast)
{ static_assert(KW >= 2 && KW <= 4 && (!WIDE || KW == 4) && BK_DC < 65535 && BK_TS >= 2*BK_DC,"bucket kernel geometry");
  typedef Key<WIDE ? 3 : 2> Entry;
  extern __shared__ __align__(16) unsigned char s_raw[];
  u64 *k0    = (u64 *) s_raw;                           /* [BK_DC] key bits 127..64                    */
  u64 *k1    = k0 + BK_DC;                              /* [BK_DC] key bits 63..0 (unused when KW == 2) */
  u32 *cnt   = (u32 *) (k1 + BK_DC);                    /* [BK_DC] instances of key i                   */
  u32 *slot  = cnt + BK_DC;                             /* [BK_TS] 0 = empty, else (fp << 16) | (i + 1) */
  u32 *sbase = slot + BK_TS;                            /* [BK_GC][BK_ROW] base strings, row = warp*32 + lane */
  __shared__ u32 s_hist[SC_SMALLHIST], s_nkeys, s_ovf, s_wsum[BK_WARPS], s_woff[BK_WARPS];
  __shared__ u64 s_ebase;

  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u64 pmask = (1ull << p.pbits) - 1ull;
  Entry *ent = (Entry *) p.ent;
  u64 ndist = 0;                                        /* thread 0: distinct keys seen by this CTA     */

  for (u32 i = threadIdx.x; i < SC_SMALLHIST; i += BK_TPB) s_hist[i] = 0;
  for (u32 i = threadIdx.x; i < BK_DC; i += BK_TPB) cnt[i] = 0;

  for (long long g = blockIdx.x; g < p.nitems; g += gridDim.x)
    { const u64 r0 = p.starts[g], r1 = p.ends[g];
      if (r1 <= r0) continue;
      if (r1 - r0 > (u64) p.big)
        { if (warp == 0) spill_group(p,g,r0,r1,lane);
          continue;
        }
      u32 R = 1, rd = 0;                                /* current hash class: keys with ((h >> 20) & (R-1)) == rd */
      for (;;)
        { /* ---- clear ---- */
          { uint4 *s4 = (uint4 *) slot;
            const uint4 z = make_uint4(0,0,0,0);
#pragma unroll
            for (u32 i = 0; i < BK_TS/4/BK_TPB; i++) s4[threadIdx.x + i*BK_TPB] = z;
          }
          if (threadIdx.x == 0) { s_nkeys = 0; s_ovf = 0; }
          __syncthreads();

          /* ---- load + expand + count: warps are independent ---- */
          u32 spare = 0xffffffffu;                      /* a key index this thread allocated and has not used yet */
          for (u64 q0 = r0; q0 < r1; q0 += BK_GC)
            { if (*(volatile u32 *) &s_ovf) break;
              const u32 ns  = (u32) ((r1 - q0 < (u64) BK_GC) ? (r1 - q0) : (u64) BK_GC);
              const u32 per = (ns + BK_WARPS - 1) / BK_WARPS;           /* <= 32 super-mers for each warp */
              const u32 w0  = warp * per;
              const u32 nw_ = (w0 < ns) ? ((ns - w0 < per) ? (ns - w0) : per) : 0u;
              u32 *rows = sbase + (warp*32)*BK_ROW;
              u32 l = 0;
              if (lane < nw_)
                { const u64 sm = p.recs[q0 + w0 + lane];
                  l = sm_len(sm,p.pbits);
                  u32 d[8];
                  u32 *row = rows + lane*BK_ROW;
                  load_supermer<PAY>(p,sm,pmask,l,false,d);
                  ((uint4 *) row)[0] = make_uint4(d[0],d[1],d[2],d[3]);
                  ((uint4 *) row)[1] = make_uint4(d[4],d[5],d[6],d[7]);
                }
              /* warp prefix of the lengths: instance x of the warp belongs to the super-mer i with pre[i] <= x < pre[i] + l[i] */
              u32 incl = l;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1)
                { const u32 y = __shfl_up_sync(0xffffffffu,incl,o);
                  if ((int) lane >= o) incl += y;
                }
              const u32 T   = __shfl_sync(0xffffffffu,incl,31);
              const u32 pre = incl - l;
              __syncwarp();
              u32 before = 0;                           /* super-mers that start before the window */
              for (u32 g0 = 0; g0 < T; g0 += 32)
                { const u32 rel = pre - g0;
                  const u32 m   = __reduce_or_sync(0xffffffffu,(l != 0u && rel < 32u) ? (1u << rel) : 0u);
                  const u32 si  = before + __popc(m & (0xffffffffu >> (31u - lane))) - 1u;     /* my super-mer (lane that loaded it) */
                  before += __popc(m);
                  const u32 ps_ = __shfl_sync(0xffffffffu,pre,si & 31u);
                  if (g0 + lane < T)
                    { const u32 j = g0 + lane - ps_;
                      u32 F[KW], G[KW];
                      supermer_strands<KW>(rows + si*BK_ROW,(int) j,p.k,klast,F,G);
                      const Key<2> key = strands_canon<KW>(F,G);
                      const u32 h = bucket_hash<KW>(key);
                      if (((h >> 20) & (R-1)) == rd)
                        { const u32 fp = h & 0xffff0000u;
                          u32 x = h & (BK_TS-1);
                          for (u32 step = 0; ; step++)
                            { if (step >= BK_PROBE) { s_ovf = 1; break; }
                              u32 v = ((volatile u32 *) slot)[x];
                              if (v == 0u)
                                { if (spare == 0xffffffffu)
                                    { spare = atomicAdd(&s_nkeys,1u);
                                      if (spare >= BK_DC) { s_ovf = 1; spare = 0xffffffffu; break; }
                                    }
                                  k0[spare] = key.w[0];
                                  if (KW > 2) k1[spare] = key.w[1];
                                  __threadfence_block();
                                  const u32 old = atomicCAS(&slot[x],0u,fp | (spare + 1u));
                                  if (old == 0u) { atomicAdd(&cnt[spare],1u); spare = 0xffffffffu; break; }
                                  v = old;
                                }
                              if ((v & 0xffff0000u) == fp)
                                { const u32 ki = (v & 0xffffu) - 1u;
                                  bool eq = (((volatile u64 *) k0)[ki] == key.w[0]);
                                  if (KW > 2) eq = eq && (((volatile u64 *) k1)[ki] == key.w[1]);
                                  if (eq) { atomicAdd(&cnt[ki],1u); break; }
                                }
                              x = (x+1) & (BK_TS-1);
                            }
                        }
                    }
                }
              __syncwarp();                             /* every lane is done with the rows before the next piece overwrites them */
            }
          __syncthreads();
          const u32 nk = (s_nkeys < (u32) BK_DC) ? s_nkeys : (u32) BK_DC;
          if (s_ovf)
            { /* split the class: forget what was counted, descend to the left child (2R, rd) */
              for (u32 i = threadIdx.x; i < nk; i += BK_TPB) cnt[i] = 0;
              if (R >= BK_MAXR)
                { if (threadIdx.x == 0) atomicAdd(p.g_fail,1u);
                  __syncthreads();
                  break;
                }
              if (threadIdx.x == 0) atomicAdd(p.g_fail + 1,1u);          /* statistics: classes split after an overflow */
              R <<= 1;
              __syncthreads();
              continue;
            }

          /* ---- emit this class: histogram, max_inst, distinct entries ---- */
          u32 cv[BK_DC/BK_TPB];
          u32 mine = 0;                                 /* (# real keys << 16) | # emitted entries; both <= BK_DC/BK_TPB per thread */
#pragma unroll
          for (u32 u = 0; u < BK_DC/BK_TPB; u++)
            { const u32 i = u*BK_TPB + threadIdx.x;
              u32 c = 0;
              if (i < nk) { c = cnt[i]; cnt[i] = 0; }
              cv[u] = c;
              if (c != 0u)
                { const u32 cs = c >= 0x7fffu ? 0x7fffu : c;
                  if (cs < SC_SMALLHIST) atomicAdd(&s_hist[cs],1u);
                  else atomicAdd(p.g_hist + cs,1ull);
                  if (c >= 0x7fffu) atomicAdd(p.g_maxinst,(u64) c);
                  mine += 0x10000u + ((ent != NULL && cs >= p.ent_min) ? 1u : 0u);
                }
            }
          u32 wtot = mine;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) wtot += __shfl_xor_sync(0xffffffffu,wtot,o);
          if (lane == 0) s_wsum[warp] = wtot;
          __syncthreads();
          if (threadIdx.x == 0)
            { u32 run = 0, real = 0;
#pragma unroll
              for (int w = 0; w < BK_WARPS; w++)
                { const u32 x = s_wsum[w];
                  s_woff[w] = run; run += x & 0xffffu; real += x >> 16;
                }
              ndist += real;
              s_ebase = run ? atomicAdd(p.ent_counter,(u64) run) : 0ull;
            }
          __syncthreads();
          if (ent != NULL && (wtot & 0xffffu))
            { u64 at = s_ebase + s_woff[warp];
#pragma unroll
              for (u32 u = 0; u < BK_DC/BK_TPB; u++)
                { const u32 c  = cv[u];
                  const u32 cs = c >= 0x7fffu ? 0x7fffu : c;
                  const bool em = (c != 0u) && (cs >= p.ent_min);
                  const u32 b = __ballot_sync(0xffffffffu,em);
                  if (em)
                    { const u64 o = at + __popc(b & ((1u << lane) - 1u));
                      if (o < p.ent_cap)
                        { const u32 i = u*BK_TPB + threadIdx.x;
                          Entry e;
                          e.w[0] = k0[i];
                          if (WIDE) { e.w[1] = k1[i]; e.w[WIDE ? 2 : 1] = (u64) cs; }
                          else e.w[1] = ((KW > 2) ? k1[i] : 0ull) | (u64) cs;
                          ent[o] = e;
                        }
                    }
                  at += __popc(b);
                }
            }
          /* ---- next class: right sibling of the nearest ancestor-or-self that is a left child ---- */
          while (R > 1 && rd >= (R >> 1)) { rd -= (R >> 1); R >>= 1; }
          if (R == 1) break;
          rd += (R >> 1);
          __syncthreads();                              /* entries were read out of k0 / k1: the next class may overwrite them */
        }
      __syncthreads();
    }

  __syncthreads();
  for (u32 i = threadIdx.x; i < SC_SMALLHIST; i += BK_TPB)
    { const u32 c = s_hist[i];
      if (c) atomicAdd(p.g_hist + i,(u64) c);
    }
  if (threadIdx.x == 0 && ndist) atomicAdd(p.g_ndistinct,ndist);
}


/* ------------------------------------------------------------------------------------------------------------------ */
/*  k_bucket_count3: WARP-PRIVATE tables, duplicate super-mers counted once.
 *  A warp owns a work group (whole buckets, ~32 super-mers) from load to emit -- no block barrier in the loop.
 *    load     <= 32 super-mers at a time, one per lane, as ORIENTED zero-padded base strings in registers (load_supermer)
 *    dedupe   at 50x coverage a locus is read ~50 times and its copies sit in the same bucket: the lane hashes its string and
 *             looks it up among the group's REPRESENTATIVES (BW_RP rows of shared memory behind a small open-addressing
 *             table); a copy only adds 1 to its representative's weight, a new string becomes a representative (the
 *             reference's Supermer_Sort + weighted k-mers, MSDsort.c:458-489 + count.c:339-542, without the sort)
 *    flush    when the rows are full, and at the end of the group: only the representatives are expanded, one k-mer per lane
 *             per round (REDUX.OR + popc mapping), and each k-mer adds its representative's weight to the warp's k-mer table
 *             (fingerprinted slots, SoA keys)
 *  Bounds -> records -> base sectors of the NEXT groups are software-pipelined (loads two groups ahead, L2 prefetch one ahead).
 *  Emit, hash classes on overflow, spill of oversize groups: as in the CTA-wide kernel above.                            */

#define BW_WARPS  4
#define BW_TPB    (32*BW_WARPS)
#define BW_KC     256                /* distinct keys a warp's class may hold */
#define BW_SL     512                /* slots per warp (load <= 0.5)          */
#define BW_RP     32                 /* representative super-mers a warp holds between flushes */
#define BW_RT     (2*BW_RP)          /* slots of their table                                   */
#define BW_WBYTES ((size_t) BW_KC*8*2 + (size_t) BW_KC*4 + (size_t) BW_SL*4 + (size_t) BW_RP*BK_ROW*4 + (size_t) BW_RT*4 + (size_t) BW_RP*4 + (size_t) 32*4)
#define BW_SMEM   (BW_WBYTES*BW_WARPS)

template<int KW, bool PAY, bool WIDE>
__global__ void __launch_bounds__(BW_TPB,6) k_bucket_count3(BucketParams p, u32 klast)
{ static_assert(KW >= 2 && KW <= 4 && (!WIDE || KW == 4) && BW_KC < 65535 && BW_SL >= 2*BW_KC && BW_RP % 32 == 0 && BW_RP <= 128,"bucket kernel geometry");
  typedef Key<WIDE ? 3 : 2> Entry;
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ u32 s_hist[SC_SMALLHIST], s_nk[BW_WARPS], s_ov[BW_WARPS], s_nrep[BW_WARPS];

  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *wb = s_raw + (size_t) warp * BW_WBYTES;
  u64 *k0   = (u64 *) wb;                               /* [BW_KC] */
  u64 *k1   = k0 + BW_KC;                               /* [BW_KC] */
  u32 *cnt  = (u32 *) (k1 + BW_KC);                     /* [BW_KC] */
  u32 *slot = cnt + BW_KC;                              /* [BW_SL] */
  u32 *rows = slot + BW_SL;                             /* [BW_RP][BK_ROW] strings of the representatives */
  u32 *rt   = rows + BW_RP*BK_ROW;                      /* [BW_RT] their table: 0 = empty, else (fp : 17)(l - 1 : 6)(0)(row + 1 : 8) */
  u32 *rm   = rt + BW_RT;                               /* [BW_RP] (copies << 8) | l ; 0 = a row that lost its race */
  u32 *cp   = rm + BW_RP;                               /* [32] holders of one flush batch, compacted */
  const u64 pmask = (1ull << p.pbits) - 1ull;
  Entry *ent = (Entry *) p.ent;
  u32 ndist = 0;                                        /* warp-uniform: distinct keys this warp has seen */
  u32 nsm = 0, nex = 0;                                 /* lane-private: super-mers met / expanded        */

  for (u32 i = threadIdx.x; i < SC_SMALLHIST; i += BW_TPB) s_hist[i] = 0;
  for (u32 i = lane; i < BW_KC; i += 32) cnt[i] = 0;
  __syncthreads();

  const long long nwarps = (long long) gridDim.x * BW_WARPS;
  auto bounds = [&](long long gg, u64 &a, u64 &b)
    { if (gg < p.nitems) { a = __ldg(p.starts + gg); b = __ldg(p.ends + gg); } else { a = 0; b = 0; } };
  auto first_recs = [&](u64 a, u64 b, u64 &x, u64 &y)
    { x = (a + lane < b) ? __ldg(p.recs + a + lane) : 0ull;
      y = (a + 32 + lane < b) ? __ldg(p.recs + a + 32 + lane) : 0ull;
    };
  auto prefetch_bases = [&](u64 sm)
    { const u64 ps = sm & pmask;
      if (PAY)
        { const u32 *pw = (const u32 *) p.payload + ps;
          asm volatile("prefetch.global.L2 [%0];" :: "l"(pw));
          asm volatile("prefetch.global.L2 [%0];" :: "l"(pw + 7));
        }
      else if (p.nranks == 1)
        { const u32 *gp = p.seq + (ps >> 4);
          asm volatile("prefetch.global.L2 [%0];" :: "l"(gp));
          asm volatile("prefetch.global.L2 [%0];" :: "l"(gp + 8));
        }
    };
  volatile u32 *ovf = &s_ov[warp];
  u32 spare = 0xffffffffu;                              /* a key index this lane allocated and has not used yet */
  u32 R = 1, rd = 0;                                    /* current hash class: keys with ((h >> 20) & (R-1)) == rd */

  /* expand the representatives held now (weights final), count their k-mers, empty the rows */
  auto flush = [&]()
    { __syncwarp();
      const u32 nrr = ((volatile u32 *) s_nrep)[warp];
      const u32 nr  = nrr < (u32) BW_RP ? nrr : (u32) BW_RP;
      for (u32 b0 = 0; b0 < nr && !*ovf; b0 += 32)
        { const u32 me_ = (b0 + lane < nr) ? rm[b0 + lane] : 0u;
          const u32 w = me_ >> 8;
          const u32 l = w ? (me_ & 0xffu) : 0u;
          if (R == 1 && l) nex++;
          u32 incl = l;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1)
            { const u32 y = __shfl_up_sync(0xffffffffu,incl,o);
              if ((int) lane >= o) incl += y;
            }
          const u32 T   = __shfl_sync(0xffffffffu,incl,31);
          const u32 pre = incl - l;
          { const u32 hmask = __ballot_sync(0xffffffffu,l != 0u);
            if (l != 0u) cp[__popc(hmask & ((1u << lane) - 1u))] = (pre << 16) | ((b0 + lane) << 8) | 0u;
          }
          __syncwarp();
          u32 before = 0;
          for (u32 g0 = 0; g0 < T; g0 += 32)
            { const u32 rel = pre - g0;
              const u32 m   = __reduce_or_sync(0xffffffffu,(l != 0u && rel < 32u) ? (1u << rel) : 0u);
              const u32 ri  = before + __popc(m & (0xffffffffu >> (31u - lane))) - 1u;     /* my super-mer: the ri-th holder of the batch */
              before += __popc(m);
              if (g0 + lane < T)
                { const u32 hc_ = cp[ri & 31u];
                  const u32 si = (hc_ >> 8) & 0xffu;                                   /* its row */
                  const u32 wsi = rm[si] >> 8;
                  const u32 j = g0 + lane - (hc_ >> 16);
                  u32 F[KW], G[KW];
                  supermer_strands<KW>(rows + si*BK_ROW,(int) j,p.k,klast,F,G);
                  const Key<2> key = strands_canon<KW>(F,G);
                  const u32 h = bucket_hash<KW>(key);
                  if (((h >> 20) & (R-1)) == rd)
                    { const u32 fp = h & 0xffff0000u;
                      u32 x = h & (BW_SL-1);
                      for (u32 step = 0; ; step++)
                        { if (step >= BK_PROBE) { *ovf = 1; break; }
                          u32 v = ((volatile u32 *) slot)[x];
                          if (v == 0u)
                            { if (spare == 0xffffffffu)
                                { spare = atomicAdd(&s_nk[warp],1u);
                                  if (spare >= BW_KC) { *ovf = 1; spare = 0xffffffffu; break; }
                                }
                              k0[spare] = key.w[0];
                              if (KW > 2) k1[spare] = key.w[1];
                              __threadfence_block();
                              const u32 old = atomicCAS(&slot[x],0u,fp | (spare + 1u));
                              if (old == 0u) { atomicAdd(&cnt[spare],wsi); spare = 0xffffffffu; break; }
                              v = old;
                            }
                          if ((v & 0xffff0000u) == fp)
                            { const u32 ki = (v & 0xffffu) - 1u;
                              bool eq = (((volatile u64 *) k0)[ki] == key.w[0]);
                              if (KW > 2) eq = eq && (((volatile u64 *) k1)[ki] == key.w[1]);
                              if (eq) { atomicAdd(&cnt[ki],wsi); break; }
                            }
                          x = (x+1) & (BW_SL-1);
                        }
                    }
                }
            }
          __syncwarp();
        }
      for (u32 i = lane; i < (u32) BW_RT; i += 32) rt[i] = 0;
      if (lane == 0) s_nrep[warp] = 0;
      __syncwarp();
    };

  long long g = (long long) blockIdx.x * BW_WARPS + warp;
  u64 r0 = 0, r1 = 0, r0n, r1n, r0nn, r1nn, sa = 0, sb = 0, san, sbn;
  u64 sann_ = 0, sbnn_ = 0, r0x_ = 0, r1x_ = 0;
  bounds(g,r0,r1); bounds(g + nwarps,r0n,r1n); bounds(g + 2*nwarps,r0nn,r1nn);
  first_recs(r0,r1,sa,sb); first_recs(r0n,r1n,san,sbn);
  for ( ; g < p.nitems; g += nwarps, r0 = r0n, r1 = r1n, r0n = r0nn, r1n = r1nn, sa = san, sb = sbn, san = sann_, sbn = sbnn_, r0nn = r0x_, r1nn = r1x_)
    { if (r0n + lane < r1n && r1n - r0n <= (u64) p.big) prefetch_bases(san);
      if (r0n + 32 + lane < r1n && r1n - r0n <= (u64) p.big) prefetch_bases(sbn);
      first_recs(r0nn,r1nn,sann_,sbnn_);
      bounds(g + 3*nwarps,r0x_,r1x_);
      if (r1 <= r0) continue;
      if (r1 - r0 > (u64) p.big) { spill_group(p,g,r0,r1,lane); continue; }
      R = 1; rd = 0;
      for (;;)
        { { uint4 *s4 = (uint4 *) slot;
            const uint4 z = make_uint4(0,0,0,0);
#pragma unroll
            for (u32 i = 0; i < BW_SL/4/32; i++) s4[lane + i*32] = z;
          }
          for (u32 i = lane; i < (u32) BW_RT; i += 32) rt[i] = 0;
          if (lane == 0) { s_nk[warp] = 0; s_ov[warp] = 0; s_nrep[warp] = 0; }
          spare = 0xffffffffu;
          __syncwarp();

          for (u64 q0 = r0; q0 < r1; q0 += 32)
            { if (*ovf) break;
              const u32 ns = (u32) ((r1 - q0 < 32ull) ? (r1 - q0) : 32ull);
              u32 l = 0, hs = 0;
              u32 d[8];
              bool pending = lane < ns;
              if (pending)
                { const u64 sm = (q0 == r0) ? sa : ((q0 == r0 + 32) ? sb : p.recs[q0 + lane]);
                  l = sm_len(sm,p.pbits);
                  load_supermer<PAY>(p,sm,pmask,l,true,d);
                  hs = d[0] * 0x9E3779B1u + d[1] * 0x85EBCA77u + d[2] * 0xC2B2AE3Du + d[3] * 0x27D4EB2Fu
                     + d[4] * 0x165667B1u + d[5] * 0xD3A2646Du + d[6] * 0xFD7046C5u + d[7] * 0xB55A4F09u + l * 0x2545F491u;
                  hs ^= hs >> 15; hs *= 0x2C1B3C6Du; hs ^= hs >> 13;
                  if (R == 1) nsm++;
                }
              /* ---- dedupe against the group's representatives; a string that finds the rows full waits for a flush ---- */
              for (;;)                                    /* every flush empties the rows, so at least the winners of the next try get in */
                { if (pending)
                    { const u32 tag = (hs & 0xffff8000u) | ((l - 1u) << 9);           /* fingerprint and length: both must agree */
                      u32 x = hs & (BW_RT-1);
                      for (;;)
                        { u32 v = ((volatile u32 *) rt)[x];
                          if (v == 0u)
                            { const u32 idx = atomicAdd(&s_nrep[warp],1u);
                              if (idx >= (u32) BW_RP) break;                           /* rows full: stay pending */
                              uint4 *r4 = (uint4 *) (rows + idx*BK_ROW);
                              r4[0] = make_uint4(d[0],d[1],d[2],d[3]);
                              r4[1] = make_uint4(d[4],d[5],d[6],d[7]);
                              rm[idx] = (1u << 8) | l;
                              __threadfence_block();
                              const u32 old = atomicCAS(&rt[x],0u,tag | (idx + 1u));
                              if (old == 0u) { pending = false; break; }               /* first holder of this string */
                              rm[idx] = 0;                                             /* lost the race: the row stays unused */
                              v = old;
                            }
                          if ((v & 0xfffffe00u) == tag)
                            { const u32 o = (v & 0xffu) - 1u;
                              const volatile u32 *rr = rows + o*BK_ROW;
                              if (((rr[0] ^ d[0]) | (rr[1] ^ d[1]) | (rr[2] ^ d[2]) | (rr[3] ^ d[3]) | (rr[4] ^ d[4]) | (rr[5] ^ d[5]) | (rr[6] ^ d[6]) | (rr[7] ^ d[7])) == 0u)
                                { atomicAdd(&rm[o],1u << 8); pending = false; break; }
                            }
                          x = (x+1) & (BW_RT-1);
                        }
                    }
                  if (!__any_sync(0xffffffffu,pending)) break;
                  flush();
                  if (*ovf) break;
                }
            }
          if (!*ovf) flush();
          __syncwarp();
          const u32 nkr = ((volatile u32 *) s_nk)[warp];
          const u32 nk  = (nkr < (u32) BW_KC) ? nkr : (u32) BW_KC;
          if (*ovf)
            { for (u32 i = lane; i < nk; i += 32) cnt[i] = 0;
              if (R >= BK_MAXR)
                { if (lane == 0) atomicAdd(p.g_fail,1u);
                  __syncwarp();
                  break;
                }
              if (lane == 0) atomicAdd(p.g_fail + 1,1u);
              R <<= 1;
              __syncwarp();
              continue;
            }

          /* emit: pass 1 histogram + totals, pass 2 entries behind one global atomic */
          u32 nemit = 0;
          for (u32 b0 = 0; b0 < nk; b0 += 32)
            { const u32 i = b0 + lane;
              const u32 c = (i < nk) ? cnt[i] : 0u;
              const u32 cs = c >= 0x7fffu ? 0x7fffu : c;
              if (c != 0u)
                { if (cs < SC_SMALLHIST) atomicAdd(&s_hist[cs],1u);
                  else atomicAdd(p.g_hist + cs,1ull);
                  if (c >= 0x7fffu) atomicAdd(p.g_maxinst,(u64) c);
                }
              ndist += __popc(__ballot_sync(0xffffffffu,c != 0u));
              nemit += __popc(__ballot_sync(0xffffffffu,c != 0u && ent != NULL && cs >= p.ent_min));
            }
          u64 at = 0;
          if (nemit)
            { if (lane == 0) at = atomicAdd(p.ent_counter,(u64) nemit);
              at = __shfl_sync(0xffffffffu,at,0);
            }
          for (u32 b0 = 0; b0 < nk; b0 += 32)
            { const u32 i = b0 + lane;
              u32 c = 0;
              if (i < nk) { c = cnt[i]; cnt[i] = 0; }
              const u32 cs = c >= 0x7fffu ? 0x7fffu : c;
              const bool em = (nemit != 0u) && (c != 0u) && (cs >= p.ent_min);
              const u32 b = __ballot_sync(0xffffffffu,em);
              if (em)
                { const u64 o = at + __popc(b & ((1u << lane) - 1u));
                  if (o < p.ent_cap)
                    { Entry e;
                      e.w[0] = k0[i];
                      if (WIDE) { e.w[1] = k1[i]; e.w[WIDE ? 2 : 1] = (u64) cs; }
                      else e.w[1] = ((KW > 2) ? k1[i] : 0ull) | (u64) cs;
                      ent[o] = e;
                    }
                }
              at += __popc(b);
            }
          while (R > 1 && rd >= (R >> 1)) { rd -= (R >> 1); R >>= 1; }
          if (R == 1) break;
          rd += (R >> 1);
          __syncwarp();
        }
      __syncwarp();
    }

  __syncthreads();
  for (u32 i = threadIdx.x; i < SC_SMALLHIST; i += BW_TPB)
    { const u32 c = s_hist[i];
      if (c) atomicAdd(p.g_hist + i,(u64) c);
    }
  if (lane == 0 && ndist) atomicAdd(p.g_ndistinct,(u64) ndist);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { nsm += __shfl_xor_sync(0xffffffffu,nsm,o); nex += __shfl_xor_sync(0xffffffffu,nex,o); }
  if (lane == 0 && p.g_stat != NULL && nsm) { atomicAdd(p.g_stat,(u64) nsm); atomicAdd(p.g_stat + 1,(u64) nex); }
}

/* ------------------------------------------------------------------------------------------------------------------ */
/*  Spilled groups: every k-mer of the listed groups is written out as a canonical Key<NW> record (the unit of the record
 *  pipeline: count.c:339-542 restated per instance, as k_scan does from the reads).  One CTA per listed group, its warps
 *  take 32 super-mers at a time; a warp reserves room for its batch with one global atomic.                            */
template<int NW, int KW, bool PAY>
__global__ void __launch_bounds__(256) k_spill_expand(BucketParams p, u32 klast, Key<NW> *out, u64 out_cap, u64 *cursor)
{ __shared__ __align__(16) u32 s_rows[8][32*BK_ROW];
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u64 pmask = (1ull << p.pbits) - 1ull;
  u32 *rows = s_rows[warp];
  const long long g = p.spill_list[blockIdx.x];
  const u64 r0 = p.starts[g], r1 = p.ends[g];
  for (u64 q0 = r0 + 32ull*warp; q0 < r1; q0 += 32ull*8)
    { const u32 ns = (u32) ((r1 - q0 < 32ull) ? (r1 - q0) : 32ull);
      u32 l = 0;
      if (lane < ns)
        { const u64 sm = p.recs[q0 + lane];
          l = sm_len(sm,p.pbits);
          u32 d[8];
          u32 *row = rows + lane*BK_ROW;
          load_supermer<PAY>(p,sm,pmask,l,false,d);
          ((uint4 *) row)[0] = make_uint4(d[0],d[1],d[2],d[3]);
          ((uint4 *) row)[1] = make_uint4(d[4],d[5],d[6],d[7]);
        }
      u32 incl = l;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
        { const u32 y = __shfl_up_sync(0xffffffffu,incl,o);
          if ((int) lane >= o) incl += y;
        }
      const u32 T   = __shfl_sync(0xffffffffu,incl,31);
      const u32 pre = incl - l;
      u64 base = 0;
      if (lane == 0) base = atomicAdd(cursor,(u64) T);
      base = __shfl_sync(0xffffffffu,base,0);
      __syncwarp();
      u32 before = 0;
      for (u32 g0 = 0; g0 < T; g0 += 32)
        { const u32 rel = pre - g0;
          const u32 m   = __reduce_or_sync(0xffffffffu,(l != 0u && rel < 32u) ? (1u << rel) : 0u);
          const u32 si  = before + __popc(m & (0xffffffffu >> (31u - lane))) - 1u;
          before += __popc(m);
          const u32 ps_ = __shfl_sync(0xffffffffu,pre,si & 31u);
          if (g0 + lane < T)
            { const u32 j = g0 + lane - ps_;
              u32 F[KW], G[KW];
              supermer_strands<KW>(rows + si*BK_ROW,(int) j,p.k,klast,F,G);
              const Key<2> key = strands_canon<KW>(F,G);
              const u64 o = base + g0 + lane;
              if (o < out_cap)
                { Key<NW> rec;
                  rec.w[0] = key.w[0];
                  if (NW > 1) rec.w[NW > 1 ? 1 : 0] = key.w[1];
                  out[o] = rec;
                }
            }
        }
      __syncwarp();
    }
}

/*  table records [kbytes key][u16 LE count] (what the record pipeline leaves) -> distinct entries appended at ent[at0 ..):
 *  the spilled k-mers rejoin the entries of the on-chip count before the key-order sort.                               */
template<int EW>
__global__ void __launch_bounds__(256) k_table_to_entries(const uint8_t *tab, u64 n, int kbytes, Key<EW> *ent, u64 at0, u64 ent_cap)
{ const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || at0 + i >= ent_cap) return;
  const uint8_t *e = tab + i * (u64) (kbytes + 2);
  u64 w[2] = { 0ull, 0ull };
  for (int b = 0; b < kbytes; b++) w[b >> 3] |= (u64) e[b] << (56 - 8*(b & 7));
  const u64 c = (u64) e[kbytes] | ((u64) e[kbytes+1] << 8);
  Key<EW> r;
  r.w[0] = w[0];
  if (EW == 3) { r.w[1] = w[1]; r.w[EW == 3 ? 2 : 1] = c; }
  else r.w[1] = w[1] | c;
  ent[at0 + i] = r;
}

}  // namespace fk
