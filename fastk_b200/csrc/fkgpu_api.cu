/*  fkgpu_api.cu -- C ABI (include/fastk_gpu.h) and host-side orchestration of the sm_100a kernels in
 *  fkgpu_kernels.cuh.  One context = one GPU = one stream.  No CPU fallback anywhere: every entry point
 *  either runs the CUDA path or returns an error.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <time.h>
#include <vector>
#include <atomic>
#include <mutex>
#include <shared_mutex>
#include <algorithm>

#include "fastk_gpu.h"
#include "fkgpu_kernels.cuh"
#include "fkgpu_bucket.cuh"

using namespace fk;

static thread_local char g_err[1024] = "";

static int set_err(int code, const char *fmt, ...)
{ va_list ap;
  va_start(ap,fmt);
  vsnprintf(g_err,sizeof(g_err),fmt,ap);
  va_end(ap);
  return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return set_err(FKGPU_E_CUDA,"%s failed at %s:%d: %s",#call,__FILE__,__LINE__,cudaGetErrorString(e_)); } while (0)

struct DevBuf
  { void *p = nullptr; size_t cap = 0;
    int ensure(size_t bytes)
    { if (bytes <= cap) return 0;
      if (p) cudaFree(p);
      p = nullptr; cap = 0;
      size_t want = bytes + (bytes >> 4) + 4096;
      if (cudaMalloc(&p,want) != cudaSuccess)
        { cudaGetLastError();
          if (cudaMalloc(&p,bytes) != cudaSuccess) { cudaGetLastError(); p = nullptr; return 1; }
          want = bytes;
        }
      cap = want;
      return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  };

struct PinBuf
  { void *p = nullptr; size_t cap = 0;
    int ensure(size_t bytes)
    { if (bytes <= cap) return 0;
      if (p) cudaFreeHost(p);
      p = nullptr; cap = 0;
      if (cudaMallocHost(&p,bytes + 4096) != cudaSuccess) { cudaGetLastError(); p = nullptr; return 1; }
      cap = bytes + 4096;
      return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  };

#include "fkgpu_multi.cuh"

#define FKGPU_D2H_CHUNKS 8          /* the table leaves the device in this many key ranges, each copied while the next is sorted */
#define CHUNK_BYTES (8u << 20)         /* pinned staging chunk per ingest thread; also the granule of the streamed pack + scan */

struct SuperGeom { int k, m, w, p2, bbits, P1, P2, pbits; };

struct TidState
  { char   *pin = nullptr;            /* pinned staging chunk                       */
    size_t  fill = 0;
    cudaEvent_t done = nullptr;       /* last async copy out of pin                 */
    bool    inflight = false;
    std::vector<std::pair<long long,long long> > chunks;   /* (device offset, bytes)            */
    std::vector<long long> rstart;    /* device position of every read start        */
    std::vector<int>       rlen;      /* read length (without terminator)           */
    std::vector<char>      rcont;     /* 1 = this read continues the previous one (rem > 0 carry) */
    int     carry = 0;                /* previous block ended with rem > 0          */
    /* direct path (the caller's DATA_BLOCK is page-locked): blocks are DMA'd straight into a device region of chunk_bytes */
    cudaStream_t str = nullptr;       /* this tid's copy stream                      */
    cudaEvent_t  rdone = nullptr;     /* region complete (copies + tail memset)      */
    long long    reg_off = -1;        /* open region, -1 = none                      */
    size_t       reg_fill = 0;
  };

struct fkgpu_ctx
  { fkgpu_config cfg;
    int  NW, kbytes;
    cudaStream_t st = nullptr, cst = nullptr;
    cudaEvent_t  ev[2*FKGPU_NSTAGES + 4];
    float        ms[FKGPU_NSTAGES];
    float        ms_bank[FKGPU_NSTAGES];
    double       bytes[FKGPU_NSTAGES];
    bool         used[FKGPU_NSTAGES];
    long long    launches = 0;
    int          sms = 148;

    /* ingest */
    std::mutex   mu;
    std::shared_mutex buf_mu;         /* shared: a direct copy into c->ascii is in flight; exclusive: c->ascii is being moved */
    std::vector<TidState> tids;
    DevBuf       ascii;               /* device copy of the ingested blocks           */
    long long    ascii_used = 0;
    long long    nreads = 0, nbases = 0;
    bool         finished = false;
    /* streamed front end (cfg.reserve_bases > 0): every staging chunk is packed -- and, on the super-mer path, scanned
       into super-mer records -- as soon as its host-to-device copy lands, overlapping the ingest                       */
    size_t       chunk_bytes = CHUNK_BYTES;   /* FKGPU_CHUNK_BYTES overrides (tests force many small chunks) */
    std::atomic<bool> stream_started{false}, stream_on{false}, stream_scan{false};   /* written under mu, polled without it */
    long long    stream_cap = 0;      /* positions the device read buffers (ascii / seq / val) were sized for: the reservation
                                         plus room for the zero gaps that end chunks and direct regions               */
    long long    stream_nub = 0;      /* bases the record buffers (super-mers, entries) were sized for                */
    SuperGeom    sgeom;

    /* device working set */
    DevBuf ctah, ctao, seq, val, bufA, bufB, scnt, hist1, off1, cur1, off2, gstart, eall, epass, poff, bsum, ghist, misc, table;
    DevBuf segs, child, pcl, sub_s, sub_e, sub_f, sub_ea, sub_ep, sub_off, sub_par, sub_base, rstart_d, prof_d;
    PinBuf h_table, h_misc, h_prof, h_poff;
    PinBuf      *h_out = nullptr;      /* where the table of the current sort goes (a run buffer of a multi-round count), default h_table */
    std::vector<PinBuf *> h_runs;      /* pinned run buffers of multi-round counts (kept for re-use)                      */
    std::vector<int64_t>  run_n;       /* result: records of every sorted run                                             */
    std::vector<const uint8_t *> run_p;/* result: host pointer of every sorted run                                        */
    DevBuf       spillA, spillB, spill_list;   /* oversize bucket groups: their k-mers as records, the list of the groups */
    DevBuf       bufC, roff1, l1k;     /* multi-round count: second entry buffer, record level-1 offsets, k-mers per level-1 bucket */
    cudaEvent_t  ev_sorted[FKGPU_D2H_CHUNKS], ev_d2h = nullptr;
    bool         d2h_pending = false;  /* a table copy on the copy stream still reads c->table                            */
    long long    st_rounds = 0, st_split = 0, st_spill = 0, st_spill_acc = 0, st_expanded = 0, st_exp_acc = 0;
    int64_t h_hist[FKGPU_HIST_BINS];

    /* profile lookup table built by finish when cfg.do_profile */
    DevBuf eprof, qoff, pkeys, pcnts, pidx, praw, pout, psrc, pdst, plen;
    DevBuf phash, phcnt, vtsrc, vtplo, vtpend;     /* hash form of the lookup table; virtual tiles of an ordered profile run */
    unsigned long long ph_nbuckets = 0; int ph_wide = 0; bool ph_on = false;
    cudaEvent_t ev_prof = nullptr;
    long long ptab_n = 0; int ptab_B = 0; long long last_npos = 0;
    bool rel_table = false;     /* -p:<table>: the lookup table was loaded from an existing k-mer table, nothing is counted */
    MultiState *mg = nullptr;   /* multi-GPU: communicator + exchange buffers (fkgpu_comm_init) */
    int weighted = 0;      /* the records being sorted are distinct (key|count) entries of the super-mer path */
    int res_nw = 2;        /* words per key of the staged / profile-table keys of the last result            */
    int last_path = 0;     /* 0 = record path, 1 = super-mer path                                            */
    long long st_super = 0, st_ent = 0, st_groups = 0;
    long long last_ndist = 0;
  };

/* ------------------------------------------------------------------------------------------------ */

extern "C" int fkgpu_device_count(void)
{ int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" const char *fkgpu_last_error(void) { return g_err; }

extern "C" int fkgpu_record_bytes(int kmer) { return (kmer <= 32) ? 8 : 16; }

extern "C" void fkgpu_packed_words(int64_t npos, int64_t *seq_words, int64_t *val_words)
{ int64_t vw = (npos + 31) / 32;
  if (val_words) *val_words = vw + FKGPU_PACK_PAD;
  if (seq_words) *seq_words = 2*vw + FKGPU_PACK_PAD;
}

extern "C" int fkgpu_create(const fkgpu_config *cfg, fkgpu_ctx **out)
{ if (cfg == NULL || out == NULL) return set_err(FKGPU_E_ARG,"fkgpu_create: NULL argument");
  *out = NULL;
  if (cfg->kmer < 1) return set_err(FKGPU_E_ARG,"fkgpu_create: k = %d must be positive",cfg->kmer);
  if (cfg->kmer > FKGPU_MAX_K) return set_err(FKGPU_E_UNSUPPORTED,"fkgpu_create: k = %d > %d not supported",cfg->kmer,FKGPU_MAX_K);
  if (cfg->do_table < 0 || cfg->bc_prefix < 0 || cfg->nthreads < 0)
    return set_err(FKGPU_E_ARG,"fkgpu_create: negative option");
  int ndev = fkgpu_device_count();
  if (ndev <= 0) return set_err(FKGPU_E_NODEVICE,"fkgpu_create: no CUDA device available (this library has no CPU path)");
  if (cfg->device < 0 || cfg->device >= ndev) return set_err(FKGPU_E_ARG,"fkgpu_create: device %d out of range [0,%d)",cfg->device,ndev);
  CU(cudaSetDevice(cfg->device));
  fkgpu_ctx *c = new (std::nothrow) fkgpu_ctx();
  if (c == NULL) return set_err(FKGPU_E_NOMEM,"fkgpu_create: out of host memory");
  c->cfg = *cfg;
  if (c->cfg.nthreads < 1) c->cfg.nthreads = 1;
  c->NW = (cfg->kmer <= 32) ? 1 : 2;
  c->kbytes = (2*cfg->kmer + 7) >> 3;
  c->tids.resize(c->cfg.nthreads);
  { const char *e = getenv("FKGPU_CHUNK_BYTES");
    if (e && atoll(e) >= 4096) c->chunk_bytes = ((size_t) atoll(e) + 63) & ~(size_t) 63;
  }
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop,cfg->device));
  c->sms = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&c->st,cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->cst,cudaStreamNonBlocking));
  for (auto &e : c->ev) CU(cudaEventCreate(&e));
  for (auto &e : c->ev_sorted) CU(cudaEventCreateWithFlags(&e,cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_d2h,cudaEventDisableTiming));
  memset(c->ms,0,sizeof(c->ms));
  memset(c->ms_bank,0,sizeof(c->ms_bank));
  memset(c->bytes,0,sizeof(c->bytes));
  memset(c->used,0,sizeof(c->used));
  *out = c;
  return FKGPU_OK;
}

static void mg_destroy(fkgpu_ctx *c);

extern "C" void fkgpu_destroy(fkgpu_ctx *c)
{ if (c == NULL) return;
  cudaSetDevice(c->cfg.device);
  cudaDeviceSynchronize();
  for (auto &t : c->tids)
    { if (t.pin) cudaFreeHost(t.pin);
      if (t.done) cudaEventDestroy(t.done);
      if (t.rdone) cudaEventDestroy(t.rdone);
      if (t.str) cudaStreamDestroy(t.str);
    }
  DevBuf *bufs[] = { &c->ctah,&c->ctao,&c->ascii,&c->seq,&c->val,&c->bufA,&c->bufB,&c->scnt,&c->hist1,&c->off1,&c->cur1,&c->off2,&c->gstart,
                     &c->eall,&c->epass,&c->poff,&c->bsum,&c->ghist,&c->misc,&c->table,&c->segs,&c->child,&c->pcl,
                     &c->sub_s,&c->sub_e,&c->sub_f,&c->sub_ea,&c->sub_ep,&c->sub_off,&c->sub_par,&c->sub_base,
                     &c->rstart_d,&c->prof_d,&c->eprof,&c->qoff,&c->pkeys,&c->pcnts,&c->pidx,&c->praw,&c->pout,&c->psrc,&c->pdst,&c->plen,
                     &c->phash,&c->phcnt,&c->vtsrc,&c->vtplo,&c->vtpend };
  for (auto b : bufs) b->release();
  if (c->ev_prof) cudaEventDestroy(c->ev_prof);
  c->h_table.release(); c->h_misc.release(); c->h_prof.release(); c->h_poff.release();
  if (c->mg) mg_destroy(c);
  for (auto r : c->h_runs) { r->release(); delete r; }
  c->bufC.release(); c->roff1.release(); c->l1k.release(); c->spillA.release(); c->spillB.release(); c->spill_list.release();
  for (auto &e : c->ev_sorted) cudaEventDestroy(e);
  if (c->ev_d2h) cudaEventDestroy(c->ev_d2h);
  for (auto &e : c->ev) cudaEventDestroy(e);
  if (c->st) cudaStreamDestroy(c->st);
  if (c->cst) cudaStreamDestroy(c->cst);
  delete c;
}

extern "C" int fkgpu_reset(fkgpu_ctx *c)
{ if (c == NULL) return set_err(FKGPU_E_ARG,"fkgpu_reset: NULL context");
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaStreamSynchronize(c->cst));
  CU(cudaStreamSynchronize(c->st));
  for (auto &t : c->tids)
    { t.fill = 0; t.inflight = false; t.chunks.clear(); t.rstart.clear(); t.rlen.clear(); t.rcont.clear(); t.carry = 0;
      if (t.str) CU(cudaStreamSynchronize(t.str));
      t.reg_off = -1; t.reg_fill = 0;
    }
  c->ascii_used = 0; c->nreads = 0; c->nbases = 0; c->finished = false;
  c->stream_started = false; c->stream_on = false; c->stream_scan = false;
  return FKGPU_OK;
}

extern "C" int fkgpu_read_counts(fkgpu_ctx *c, int64_t *per_tid)
{ if (c == NULL || per_tid == NULL) return set_err(FKGPU_E_ARG,"fkgpu_read_counts: NULL argument");
  for (size_t t = 0; t < c->tids.size(); t++)
    { long long n = 0;
      for (char x : c->tids[t].rcont) n += (x == 0);
      per_tid[t] = n;
    }
  return FKGPU_OK;
}

extern "C" int fkgpu_last_path(fkgpu_ctx *c) { return c ? c->last_path : 0; }

extern "C" int fkgpu_last_stats(fkgpu_ctx *c, int64_t *v)
{ if (c == NULL || v == NULL) return set_err(FKGPU_E_ARG,"fkgpu_last_stats: NULL argument");
  v[0] = c->last_path; v[1] = c->st_super; v[2] = c->st_ent; v[3] = c->st_groups;
  v[4] = c->st_rounds; v[5] = c->st_split; v[6] = c->st_spill; v[7] = c->st_expanded;
  return FKGPU_OK;
}

extern "C" int64_t fkgpu_launch_count(fkgpu_ctx *c) { return c ? c->launches : 0; }

extern "C" int fkgpu_stage_times(fkgpu_ctx *c, float *ms, double *bytes)
{ if (c == NULL) return set_err(FKGPU_E_ARG,"fkgpu_stage_times: NULL context");
  for (int i = 0; i < FKGPU_NSTAGES; i++)
    { if (ms) ms[i] = c->ms[i];
      if (bytes) bytes[i] = c->bytes[i];
    }
  return FKGPU_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/*  ingest                                                                                           */

static int ascii_reserve(fkgpu_ctx *c, long long need)      /* c->mu held */
{ if ((size_t) need + 64 <= c->ascii.cap) return 0;
  size_t want = std::max((size_t) need + 64, std::max(c->ascii.cap*2,(size_t) 256 << 20));
  if (c->cfg.reserve_bases > 0)
    want = std::max(want,(size_t) (c->cfg.reserve_bases + c->cfg.reserve_bases/50 + (1 << 20)));
  void *np = nullptr;
  if (cudaMalloc(&np,want) != cudaSuccess) { cudaGetLastError(); return 1; }
  std::unique_lock<std::shared_mutex> moving(c->buf_mu);      /* no direct copy may be in flight while the buffer moves */
  if (c->ascii.p)
    { cudaStreamSynchronize(c->cst);
      cudaMemcpy(np,c->ascii.p,(size_t) c->ascii_used,cudaMemcpyDeviceToDevice);
      cudaFree(c->ascii.p);
    }
  c->ascii.p = np; c->ascii.cap = want;
  return 0;
}

static int stream_begin(fkgpu_ctx *c);
static int stream_chunk(fkgpu_ctx *c, long long off, long long len, long long scan_len);

static int flush_tid(fkgpu_ctx *c, TidState &t)
{ if (t.fill == 0) return FKGPU_OK;
  /* chunks start on multiples of 64 positions (whole seq / val words per chunk); the gap is zero = invalid positions */
  const size_t padded = (t.fill + 63) & ~(size_t) 63;
  memset(t.pin + t.fill,0,padded - t.fill);
  long long off;
  { std::lock_guard<std::mutex> lk(c->mu);
    if (c->stream_on && c->ascii_used + (long long) padded > c->stream_cap)
      { /* more reads than reserved: the buffers must grow, so the rest is packed and scanned at finish instead */
        CU(cudaStreamSynchronize(c->st));
        c->stream_on = false; c->stream_scan = false;
      }
    if (ascii_reserve(c,c->ascii_used + (long long) padded))
      return set_err(FKGPU_E_NOMEM,"fkgpu_ingest: cannot grow the device read buffer to %lld bytes",c->ascii_used + (long long) padded);
    off = c->ascii_used;
    c->ascii_used += (long long) padded;
    CU(cudaMemcpyAsync((char *) c->ascii.p + off,t.pin,padded,cudaMemcpyHostToDevice,c->cst));
    CU(cudaEventRecord(t.done,c->cst));
    if (c->stream_on)
      { CU(cudaStreamWaitEvent(c->st,t.done,0));
        int rc = stream_chunk(c,off,(long long) padded,(long long) padded);
        if (rc) return rc;
      }
  }
  t.inflight = true;
  t.chunks.push_back(std::make_pair(off,(long long) t.fill));
  /* rebase the read starts of this chunk (they were recorded chunk-relative, tagged negative) */
  for (size_t i = t.rstart.size(); i-- > 0; )
    { if (t.rstart[i] >= 0) break;
      t.rstart[i] = off + (-(t.rstart[i]) - 1);
    }
  t.fill = 0;
  return FKGPU_OK;
}

/*  Direct path: the caller's block is page-locked (cudaMallocHost / cudaHostRegister), so it is DMA'd straight to its
 *  place in the device read buffer -- no staging memcpy on the host.  Each tid fills device regions of chunk_bytes; a
 *  complete region gets its tail zeroed (invalid positions) and is packed + scanned like a staged chunk.  The copy is
 *  waited for before returning: the caller reuses the block immediately (io.c:552,565).                            */
static int close_region(fkgpu_ctx *c, TidState &t)
{ if (t.reg_off < 0) return FKGPU_OK;
  const size_t tail = c->chunk_bytes - t.reg_fill;
  { std::shared_lock<std::shared_mutex> stable(c->buf_mu);
    if (tail) CU(cudaMemsetAsync((char *) c->ascii.p + t.reg_off + t.reg_fill,0,tail,t.str));
    CU(cudaEventRecord(t.rdone,t.str));
    CU(cudaStreamSynchronize(t.str));
  }
  { std::lock_guard<std::mutex> lk(c->mu);
    if (c->stream_on)
      { int rc = stream_chunk(c,t.reg_off,(long long) c->chunk_bytes,(long long) ((t.reg_fill + 63) & ~(size_t) 63));
        if (rc) return rc;
      }
  }
  t.chunks.push_back(std::make_pair(t.reg_off,(long long) t.reg_fill));
  t.reg_off = -1; t.reg_fill = 0;
  return FKGPU_OK;
}

/*  -> 1 if the block was taken (its device position in *pos0), 0 if the caller must use the staging path */
static int direct_ingest(fkgpu_ctx *c, TidState &t, const char *src, size_t len, long long *pos0)
{ static int off = -1;
  if (off < 0) { const char *e = getenv("FKGPU_NODIRECT"); off = (e && atoi(e)) ? 1 : 0; }
  if (off || !c->stream_on) return 0;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr,src) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (attr.type != cudaMemoryTypeHost) return 0;
  if (t.str == nullptr)
    { if (cudaStreamCreateWithFlags(&t.str,cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&t.rdone,cudaEventDisableTiming) != cudaSuccess)
        { cudaGetLastError(); return 0; }
    }
  if (t.reg_off < 0 || t.reg_fill + len > c->chunk_bytes)
    { int rc = close_region(c,t);
      if (rc) return rc;
      std::lock_guard<std::mutex> lk(c->mu);
      if (!c->stream_on || c->ascii_used + (long long) c->chunk_bytes > c->stream_cap) return 0;
      t.reg_off = c->ascii_used;
      c->ascii_used += (long long) c->chunk_bytes;
      t.reg_fill = 0;
    }
  { std::shared_lock<std::shared_mutex> stable(c->buf_mu);
    if (cudaMemcpyAsync((char *) c->ascii.p + t.reg_off + t.reg_fill,src,len,cudaMemcpyHostToDevice,t.str) != cudaSuccess
        || cudaStreamSynchronize(t.str) != cudaSuccess)
      return set_err(FKGPU_E_CUDA,"fkgpu_ingest: direct copy failed: %s",cudaGetErrorString(cudaGetLastError()));
  }
  *pos0 = t.reg_off + (long long) t.reg_fill;
  t.reg_fill += len;
  return 1;
}

extern "C" int fkgpu_ingest(fkgpu_ctx *c, int tid, const char *bases, const int32_t *boff, int32_t nreads, int32_t rem)
{ if (c == NULL || (nreads > 0 && (bases == NULL || boff == NULL)))
    return set_err(FKGPU_E_ARG,"fkgpu_ingest: NULL argument");
  if (tid < 0 || tid >= (int) c->tids.size()) return set_err(FKGPU_E_ARG,"fkgpu_ingest: tid %d out of range [0,%d)",tid,(int) c->tids.size());
  if (c->finished) return set_err(FKGPU_E_STATE,"fkgpu_ingest: called after fkgpu_finish (use fkgpu_reset)");
  if (nreads <= 0) return FKGPU_OK;
  CU(cudaSetDevice(c->cfg.device));
  TidState &t = c->tids[tid];
  if (!c->stream_started)
    { std::lock_guard<std::mutex> lk(c->mu);
      if (!c->stream_started)
        { int rc = stream_begin(c);
          if (rc) return rc;
          c->stream_started = true;                 /* published last: the buffers and flags above are final */
        }
    }
  if (t.pin == nullptr)
    { if (cudaMallocHost((void **) &t.pin,c->chunk_bytes + 64) != cudaSuccess)
        { cudaGetLastError(); return set_err(FKGPU_E_NOMEM,"fkgpu_ingest: cannot allocate pinned staging"); }
      CU(cudaEventCreateWithFlags(&t.done,cudaEventDisableTiming));
    }
  const size_t len = (size_t) boff[nreads] - (size_t) boff[0];
  if (len > c->chunk_bytes) return set_err(FKGPU_E_ARG,"fkgpu_ingest: block of %zu bytes exceeds the %zu byte staging chunk",len,c->chunk_bytes);
  long long dpos = -1;
  { int took = 0;
    if (c->stream_on)
      { if (t.fill > 0 && t.reg_off < 0)                 /* keep this tid's pieces in arrival order */
          { cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr,bases) == cudaSuccess && attr.type == cudaMemoryTypeHost)
              { int rc = flush_tid(c,t);
                if (rc) return rc;
              }
            else cudaGetLastError();
          }
        if (t.fill == 0)
          { took = direct_ingest(c,t,bases + boff[0],len,&dpos);
            if (took < 0) return took;
          }
      }
    if (!took)
      { dpos = -1;
        int rc = close_region(c,t);                      /* a tid that falls back to staging finishes its open region first */
        if (rc) return rc;
      }
  }
  if (dpos < 0 && t.fill + len > c->chunk_bytes)
    { int rc = flush_tid(c,t);
      if (rc) return rc;
    }
  if (dpos < 0)
    { if (t.inflight && t.fill == 0)
        { CU(cudaEventSynchronize(t.done));      /* the chunk is being reused: previous copy must be out */
          t.inflight = false;
        }
      memcpy(t.pin + t.fill,bases + boff[0],len);
    }
  long long nb = 0;
  for (int i = 0; i < nreads; i++)
    { long long s = (dpos >= 0 ? dpos : (long long) t.fill) + (boff[i] - boff[0]);
      int rl = boff[i+1] - boff[i] - 1;
      t.rstart.push_back(dpos >= 0 ? s : -(s + 1));      /* staged: chunk relative, fixed up at flush */
      t.rlen.push_back(rl);
      t.rcont.push_back((i == 0 && t.carry) ? 1 : 0);
      nb += rl;
    }
  /* a continued read re-delivers its k-1 overlap; count it once (split.c:1046-1053) */
  long long nr = nreads;
  if (t.carry) { nr -= 1; nb -= (c->cfg.kmer - 1); }
  t.carry = (rem > 0);
  if (dpos < 0) t.fill += len;
  { std::lock_guard<std::mutex> lk(c->mu);
    c->nreads += nr;
    c->nbases += nb;
  }
  return FKGPU_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/*  the counting pipeline                                                                            */

struct Misc            /* small device-side scalars, one cudaMemcpy to read them all */
  { u64 maxinst, ndistinct, total_pass;
    u32 ovf_cnt, ticket;
  };

static void stage_begin(fkgpu_ctx *c, int s) { cudaEventRecord(c->ev[2*s],c->st); c->used[s] = true; }
static void stage_end  (fkgpu_ctx *c, int s) { cudaEventRecord(c->ev[2*s+1],c->st); }

static int ilog2_ceil(unsigned long long x) { int l = 0; while ((1ull << l) < x) l++; return l; }

#define KCHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
    return set_err(FKGPU_E_CUDA,"kernel launch failed at %s:%d: %s",__FILE__,__LINE__,cudaGetErrorString(e_)); c->launches++; } while (0)

/*  FKGPU_PROF=legacy: profiles by binary search over the sorted keys behind a prefix index, one u16 per position, pieces
 *  gathered afterwards (round 1 .. r2h).  Default: hash lookups, counts written straight to their place in the output, the
 *  output copied to the host in slices while later tiles are looked up.                                                */
static bool prof_legacy()
{ static int v = -1;
  if (v < 0) { const char *e = getenv("FKGPU_PROF"); v = (e && strcmp(e,"legacy") == 0) ? 1 : 0; }
  return v == 1;
}

/*  (pkeys, pcnts)[0..U) hold the distinct k-mers in key order: build what k_profile looks them up with  */
template<int NW>
static int build_profile_lookup(fkgpu_ctx *c, u64 U)
{ if (prof_legacy())
    { int B = ilog2_ceil(U + 1); if (B < 8) B = 8; if (B > 28) B = 28;
      if (c->pidx.ensure(((size_t) (1ull << B) + 2) * 8)) return set_err(FKGPU_E_NOMEM,"out of device memory (profile index of %llu k-mers)",U);
      k_build_index<NW><<<(unsigned) ((U + 1 + 255) / 256),256,0,c->st>>>((const Key<NW> *) c->pkeys.p,U,B,(u64 *) c->pidx.p); KCHECK();
      c->ptab_B = B; c->ph_on = false;
      return FKGPU_OK;
    }
  const u64 nb = (U * 7) / 16 + 64;                  /* 4 slots per bucket: load 4/7 */
  const int wide = (NW == 2 && c->cfg.kmer > 56) ? 1 : 0;
  if (c->phash.ensure((size_t) nb * 64) || (wide && c->phcnt.ensure((size_t) nb * 8)))
    return set_err(FKGPU_E_NOMEM,"out of device memory (profile hash table of %llu k-mers)",U);
  CU(cudaMemsetAsync(c->phash.p,0,(size_t) nb * 64,c->st));
  if (wide) CU(cudaMemsetAsync(c->phcnt.p,0,(size_t) nb * 8,c->st));
  if (U > 0)
    { k_hash_build<NW><<<(unsigned) ((U + 255) / 256),256,0,c->st>>>((const Key<NW> *) c->pkeys.p,(const uint16_t *) c->pcnts.p,U,(ulonglong2 *) c->phash.p,
                                                                 (uint16_t *) c->phcnt.p,nb,wide); KCHECK();
    }
  c->ph_nbuckets = nb; c->ph_wide = wide; c->ph_on = true;
  return FKGPU_OK;
}

static const u32 SC_CAP = 2048;      /* records one k_sortcount CTA can hold           */
static const u32 SC_T   = 1024;      /* group packing target (fine buckets up to SC_CAP-SC_T+1 never overflow) */

struct ScLayout { u32 tab_off, srt_off, srt2_off, total; };
static ScLayout sc_layout(int NW)
{ ScLayout L;
  u32 recb = (SC_CAP + 2) * 8 * NW;
  u32 hmax = 1; while (hmax < SC_CAP + SC_CAP/4 + 1) hmax <<= 1;
  u32 dmax = 1; while (dmax < SC_CAP) dmax <<= 1;
  L.tab_off = (recb + 127) & ~127u;
  L.srt_off = L.tab_off + hmax*4;
  L.total   = L.srt_off + dmax*8;
  L.srt2_off = dmax/2;
  return L;
}

template<int NW> static int run_large_scan(fkgpu_ctx *c, const u32 *in, long long n, u64 *out, u64 *total_dev)
{ long long nb = (n + LS_CHUNK - 1) / LS_CHUNK;
  if (nb == 0) nb = 1;
  if (c->bsum.ensure((size_t) nb * 8)) return set_err(FKGPU_E_NOMEM,"scan: out of device memory");
  k_lscan_reduce<<<(unsigned) nb,256,0,c->st>>>(in,n,(u64 *) c->bsum.p); KCHECK();
  k_lscan_top<<<1,1024,0,c->st>>>((u64 *) c->bsum.p,nb,total_dev); KCHECK();
  k_lscan_apply<<<(unsigned) nb,256,0,c->st>>>(in,n,(const u64 *) c->bsum.p,out); KCHECK();
  return FKGPU_OK;
}

/*  Everything after "records are in bufA grouped by their top P1 bits, off1[] holds the group starts".
 *  nub = upper bound on the record count used for sizing.                                           */
template<int NW>
static int count_from_level1(fkgpu_ctx *c, long long nub, int P1, int P2, int fetch_table, fkgpu_result *res, void *bufX, void *bufY)
{ typedef Key<NW> K;
  const int nb1 = 1 << P1;
  const long long m = (long long) nb1 << P2;           /* # fine buckets */
  const long long gmax = nub / SC_T + 2;
  const ScLayout L = sc_layout(NW);
  Misc *d_misc = (Misc *) c->misc.p;

  K *X = (K *) bufX, *Y = (K *) bufY;
  const u64 *offs = (const u64 *) c->off1.p;
  /* per-sort scratch scalars (total_pass, ovf_cnt, ticket): a count may run this stage more than once (oversize buckets
     through the record pipeline, then the entries; one sort per round)                                                  */
  CU(cudaMemsetAsync(&d_misc->total_pass,0,16,c->st));

  if (P2 > 0)
    { if (c->off2.ensure((size_t) (m + 1) * 8)) return set_err(FKGPU_E_NOMEM,"out of device memory (off2)");
      stage_begin(c,FKGPU_ST_L2PART);
      int grid = std::min(nb1,c->sms * 2);
      size_t sm = (size_t) REF_ST * REF_SLOT * sizeof(K) + (size_t) (1 << P2) * 4;
      CU(cudaFuncSetAttribute(k_refine<NW>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) sm));
      k_refine<NW><<<grid,REF_TPB,sm,c->st>>>(X,Y,(const u64 *) c->off1.p,nb1,P1,P2,(u64 *) c->off2.p,&d_misc->ticket); KCHECK();
      stage_end(c,FKGPU_ST_L2PART);
      offs = (const u64 *) c->off2.p;
      std::swap(X,Y);
    }
  /* now: records in X, grouped into m fine buckets with starts offs[0..m]; Y is free */

  if (c->gstart.ensure((size_t) (gmax + 2) * 8) || c->eall.ensure((size_t) gmax * 4) || c->epass.ensure((size_t) gmax * 4)
      || c->poff.ensure((size_t) (gmax + 1) * 8) || c->scnt.ensure((size_t) (nub + 2) * 4))
    return set_err(FKGPU_E_NOMEM,"out of device memory (group arrays)");
  u64 *gstart = (u64 *) c->gstart.p;
  k_fill_u64<<<(unsigned) ((gmax + 1 + 255) / 256),256,0,c->st>>>(gstart,gmax + 1,offs + m); KCHECK();
  k_groups<<<(unsigned) ((m + 1 + 255) / 256),256,0,c->st>>>(offs,m,SC_T,gstart,gmax); KCHECK();

  stage_begin(c,FKGPU_ST_SORTCOUNT);
  SortCountParams sp;
  sp.in0 = X; sp.stage0 = Y; sp.in1 = Y; sp.stage1 = X;
  sp.stage_cnt = (u32 *) c->scnt.p;
  sp.starts = gstart; sp.ends = gstart + 1; sp.flags = NULL;
  sp.e_all = (u32 *) c->eall.p; sp.e_pass = (u32 *) c->epass.p;
  sp.g_hist = (u64 *) c->ghist.p; sp.g_maxinst = &d_misc->maxinst; sp.g_ndistinct = &d_misc->ndistinct;
  if (c->segs.ensure((size_t) gmax * 4)) return set_err(FKGPU_E_NOMEM,"out of device memory (overflow list)");
  sp.ovf_cnt = &d_misc->ovf_cnt; sp.ovf_list = (u32 *) c->segs.p; sp.ovf_cap = (u32) std::min<long long>(gmax,0x7fffffffll);
  sp.cap = SC_CAP; sp.cutoff = (u32) std::max(1,c->cfg.do_table); sp.nitems = gmax; sp.item_base = 0;
  sp.tab_off = L.tab_off; sp.srt_off = L.srt_off; sp.srt2_off = L.srt2_off;
  sp.weighted = (u32) c->weighted;
  /* weighted + no profiles: every entry is distinct and already passed the cutoff in the bucket kernel, so entry i of
     the key order is table record i -- the sort kernel writes the table itself (no staging, no compaction pass)       */
  const bool direct = c->weighted && !c->cfg.do_profile && c->cfg.do_table > 0;
  const int twd = c->kbytes + 2;
  sp.direct = NULL; sp.kbytes = c->kbytes; sp.tab_bytes = L.srt_off - L.tab_off;
  if (direct)
    { if (c->table.ensure((size_t) nub * twd + 64)) return set_err(FKGPU_E_NOMEM,"out of device memory (table of %lld entries)",nub);
      sp.direct = (uint8_t *) c->table.p;
    }
  CU(cudaFuncSetAttribute(k_sortcount<NW>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) L.total));
  /* direct + fetch: entry i of the key order is table record i and their number is known (nub), so the table leaves the
     device in FKGPU_D2H_CHUNKS key ranges, each copied on the copy stream while the next range is being sorted          */
  PinBuf *hout = c->h_out ? c->h_out : &c->h_table;
  bool d2h_chunked = false;
  if (c->d2h_pending) { CU(cudaStreamWaitEvent(c->st,c->ev_d2h,0)); c->d2h_pending = false; }   /* c->table is about to be rewritten */
  if (direct && fetch_table && (size_t) nub * twd >= ((size_t) 32 << 20))
    { if (hout->ensure((size_t) nub * twd + 64)) return set_err(FKGPU_E_NOMEM,"out of pinned host memory (table)");
      long long gb[FKGPU_D2H_CHUNKS + 1];
      u64 eb[FKGPU_D2H_CHUNKS + 1];
      for (int q = 0; q <= FKGPU_D2H_CHUNKS; q++) gb[q] = gmax * q / FKGPU_D2H_CHUNKS;
      for (int q = 0; q <= FKGPU_D2H_CHUNKS; q++)
        CU(cudaMemcpyAsync(eb + q,gstart + gb[q],8,cudaMemcpyDeviceToHost,c->st));
      CU(cudaStreamSynchronize(c->st));
      for (int q = 0; q < FKGPU_D2H_CHUNKS; q++)
        { if (gb[q+1] > gb[q])
            { SortCountParams sq = sp;
              sq.starts = gstart + gb[q]; sq.ends = gstart + gb[q] + 1;
              sq.e_all = sp.e_all + gb[q]; sq.e_pass = sp.e_pass + gb[q];
              sq.nitems = gb[q+1] - gb[q]; sq.item_base = (u32) gb[q];
              k_sortcount<NW><<<(unsigned) sq.nitems,SC_TPB,L.total,c->st>>>(sq); KCHECK();
            }
          CU(cudaEventRecord(c->ev_sorted[q],c->st));
          CU(cudaStreamWaitEvent(c->cst,c->ev_sorted[q],0));
          if (eb[q+1] > eb[q])
            CU(cudaMemcpyAsync((uint8_t *) hout->p + eb[q] * twd,(const uint8_t *) c->table.p + eb[q] * twd,(size_t) (eb[q+1] - eb[q]) * twd,
                               cudaMemcpyDeviceToHost,c->cst));
        }
      CU(cudaEventRecord(c->ev_d2h,c->cst));
      c->d2h_pending = true;
      d2h_chunked = true;
    }
  else
    { k_sortcount<NW><<<(unsigned) gmax,SC_TPB,L.total,c->st>>>(sp); KCHECK(); }
  stage_end(c,FKGPU_ST_SORTCOUNT);

  /* first sync: overflow list + totals */
  Misc hm;
  CU(cudaMemcpyAsync(&hm,d_misc,sizeof(Misc),cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));

  /* ---- fallback for oversize groups (heavy repeats): host-driven MSD refinement ---------------- */
  struct Sub { u64 s, e; u32 fl; u32 parent; };
  std::vector<Sub> subs;
  std::vector<u32> sub_par_h, sub_eall_h;
  static int verbose = -1;
  if (verbose < 0) { const char *e = getenv("FKGPU_VERBOSE"); verbose = e ? atoi(e) : 0; }
  if (verbose)
    fprintf(stderr,"[fkgpu] nub=%lld P1=%d P2=%d groups=%lld oversize_groups=%u\n",nub,P1,P2,gmax,hm.ovf_cnt);
  std::vector<u64> sub_base_pass;
  if (hm.ovf_cnt > 0)
    { if (hm.ovf_cnt > sp.ovf_cap) return set_err(FKGPU_E_CUDA,"internal: overflow list truncated");
      std::vector<u32> og(hm.ovf_cnt);
      CU(cudaMemcpy(og.data(),c->segs.p,(size_t) hm.ovf_cnt * 4,cudaMemcpyDeviceToHost));
      std::sort(og.begin(),og.end());
      std::vector<Sub> work;
      for (u32 g : og)
        { u64 se[2];
          CU(cudaMemcpy(se,gstart + g,16,cudaMemcpyDeviceToHost));
          Sub w; w.s = se[0]; w.e = se[1]; w.fl = 0; w.parent = g;
          work.push_back(w);
        }
      int guard = 0;
      while (!work.empty())
        { if (++guard > 40) return set_err(FKGPU_E_CUDA,"internal: refinement did not converge");
          size_t ns = work.size();
          std::vector<u64> hs(ns), he(ns); std::vector<u32> hf(ns);
          for (size_t i = 0; i < ns; i++) { hs[i] = work[i].s; he[i] = work[i].e; hf[i] = work[i].fl; }
          if (c->sub_s.ensure(ns*8) || c->sub_e.ensure(ns*8) || c->sub_f.ensure(ns*4) || c->child.ensure(ns*257*8) || c->pcl.ensure(ns*4))
            return set_err(FKGPU_E_NOMEM,"out of device memory (refinement)");
          CU(cudaMemcpyAsync(c->sub_s.p,hs.data(),ns*8,cudaMemcpyHostToDevice,c->st));
          CU(cudaMemcpyAsync(c->sub_e.p,he.data(),ns*8,cudaMemcpyHostToDevice,c->st));
          CU(cudaMemcpyAsync(c->sub_f.p,hf.data(),ns*4,cudaMemcpyHostToDevice,c->st));
          k_autorefine<NW><<<(unsigned) ns,AR_TPB,0,c->st>>>(X,Y,(const u64 *) c->sub_s.p,(const u64 *) c->sub_e.p,
                                                            (const u32 *) c->sub_f.p,(u64 *) c->child.p,(u32 *) c->pcl.p); KCHECK();
          std::vector<u64> hc(ns*257); std::vector<u32> hp(ns);
          CU(cudaMemcpyAsync(hc.data(),c->child.p,ns*257*8,cudaMemcpyDeviceToHost,c->st));
          CU(cudaMemcpyAsync(hp.data(),c->pcl.p,ns*4,cudaMemcpyDeviceToHost,c->st));
          CU(cudaStreamSynchronize(c->st));
          std::vector<Sub> next;
          for (size_t i = 0; i < ns; i++)
            { if ((int) hp[i] >= 64*NW)
                { Sub u = work[i]; u.fl |= ITEM_UNIFORM; subs.push_back(u); continue; }
              for (int d = 0; d < 256; d++)
                { u64 s = hc[i*257+d], e = hc[i*257+d+1];
                  if (e == s) continue;
                  Sub u; u.s = s; u.e = e; u.fl = work[i].fl ^ ITEM_ALTBUF; u.parent = work[i].parent;
                  if (e - s <= SC_CAP) subs.push_back(u); else next.push_back(u);
                }
            }
          work.swap(next);
        }
      std::sort(subs.begin(),subs.end(),[](const Sub &a, const Sub &b) { return a.s < b.s; });
      size_t nsub = subs.size();
      std::vector<u64> hs(nsub), he(nsub); std::vector<u32> hf(nsub), hpar(nsub);
      for (size_t i = 0; i < nsub; i++) { hs[i] = subs[i].s; he[i] = subs[i].e; hf[i] = subs[i].fl; hpar[i] = subs[i].parent; }
      if (c->sub_s.ensure(nsub*8) || c->sub_e.ensure(nsub*8) || c->sub_f.ensure(nsub*4) || c->sub_ea.ensure(nsub*4) || c->sub_ep.ensure(nsub*4)
          || c->sub_off.ensure(nsub*8) || c->sub_par.ensure(nsub*4) || c->sub_base.ensure(nsub*8))
        return set_err(FKGPU_E_NOMEM,"out of device memory (refinement items)");
      CU(cudaMemcpyAsync(c->sub_s.p,hs.data(),nsub*8,cudaMemcpyHostToDevice,c->st));
      CU(cudaMemcpyAsync(c->sub_e.p,he.data(),nsub*8,cudaMemcpyHostToDevice,c->st));
      CU(cudaMemcpyAsync(c->sub_f.p,hf.data(),nsub*4,cudaMemcpyHostToDevice,c->st));
      CU(cudaMemcpyAsync(c->sub_par.p,hpar.data(),nsub*4,cudaMemcpyHostToDevice,c->st));
      SortCountParams sq = sp;
      sq.starts = (const u64 *) c->sub_s.p; sq.ends = (const u64 *) c->sub_e.p; sq.flags = (const u32 *) c->sub_f.p;
      sq.e_all = (u32 *) c->sub_ea.p; sq.e_pass = (u32 *) c->sub_ep.p; sq.nitems = (long long) nsub;
      k_sortcount<NW><<<(unsigned) nsub,SC_TPB,L.total,c->st>>>(sq); KCHECK();
      std::vector<u32> hep(nsub);
      sub_eall_h.resize(nsub); sub_par_h = hpar;
      CU(cudaMemcpyAsync(hep.data(),c->sub_ep.p,nsub*4,cudaMemcpyDeviceToHost,c->st));
      CU(cudaMemcpyAsync(sub_eall_h.data(),c->sub_ea.p,nsub*4,cudaMemcpyDeviceToHost,c->st));
      CU(cudaMemcpyAsync(&hm,d_misc,sizeof(Misc),cudaMemcpyDeviceToHost,c->st));
      CU(cudaStreamSynchronize(c->st));
      if (hm.ovf_cnt != (u32) og.size()) return set_err(FKGPU_E_CUDA,"internal: refinement left oversize items");
      /* per parent group: total of its sub-items goes into e_pass[parent]; each sub-item gets its base */
      std::vector<u64> hbase(nsub);
      size_t i = 0;
      while (i < nsub)
        { size_t j = i; u64 run = 0;
          while (j < nsub && hpar[j] == hpar[i]) { hbase[j] = run; run += hep[j]; j++; }
          u32 tot = (u32) run;
          CU(cudaMemcpyAsync((u32 *) c->epass.p + hpar[i],&tot,4,cudaMemcpyHostToDevice,c->st));
          CU(cudaStreamSynchronize(c->st));
          i = j;
        }
      sub_base_pass = hbase;
    }

  /* ---- -p: compact ALL distinct k-mers into the lookup table used by fkgpu_profiles ------------------- */
  c->ptab_n = 0;
  if (c->cfg.do_profile)
    { if (c->eprof.ensure((size_t) gmax * 4) || c->qoff.ensure((size_t) (gmax + 1) * 8))
        return set_err(FKGPU_E_NOMEM,"out of device memory (profile table offsets)");
      CU(cudaMemcpyAsync(c->eprof.p,c->eall.p,(size_t) gmax * 4,cudaMemcpyDeviceToDevice,c->st));
      const size_t nsub = subs.size();
      std::vector<u64> hb(nsub);
      for (size_t i = 0; i < nsub; )
        { size_t j = i; u64 run = 0;
          while (j < nsub && sub_par_h[j] == sub_par_h[i]) { hb[j] = run; run += sub_eall_h[j]; j++; }
          u32 tot = (u32) run;
          CU(cudaMemcpyAsync((u32 *) c->eprof.p + sub_par_h[i],&tot,4,cudaMemcpyHostToDevice,c->st));
          CU(cudaStreamSynchronize(c->st));
          i = j;
        }
      int rc = run_large_scan<NW>(c,(const u32 *) c->eprof.p,gmax,(u64 *) c->qoff.p,&d_misc->total_pass);
      if (rc) return rc;
      CU(cudaMemcpyAsync(&hm,d_misc,sizeof(Misc),cudaMemcpyDeviceToHost,c->st));
      CU(cudaStreamSynchronize(c->st));
      const u64 U = hm.total_pass;
      /* about one key per index slot: a lookup is the slot's two bounds (one sector), ~one key, its count (ncu r2: with 3.5 keys
         per slot k_profile moved 294 B of DRAM per lookup at 75 % of the HBM peak)                                          */
      if (c->pkeys.ensure((size_t) (U + 1) * sizeof(K)) || c->pcnts.ensure((size_t) (U + 1) * 2))
        return set_err(FKGPU_E_NOMEM,"out of device memory (profile lookup table of %llu k-mers)",U);
      CompactParams cp;
      cp.stage0 = Y; cp.stage1 = X; cp.stage_cnt = (const u32 *) c->scnt.p;
      cp.starts = gstart; cp.flags = NULL; cp.e_all = (const u32 *) c->eall.p; cp.out_off = (const u64 *) c->qoff.p;
      cp.out = NULL; cp.nitems = gmax; cp.cutoff = 0; cp.kbytes = c->kbytes;
      typedef Key<(NW == 3) ? 2 : NW> PK;            /* key words only: the third word of a wide entry held the count */
      k_compact_keys<NW><<<c->sms * 8,256,0,c->st>>>(cp,(PK *) c->pkeys.p,(uint16_t *) c->pcnts.p); KCHECK();
      if (nsub > 0)
        { if (c->sub_base.ensure(nsub*8) || c->sub_off.ensure(nsub*8)) return set_err(FKGPU_E_NOMEM,"out of device memory");
          CU(cudaMemcpyAsync(c->sub_base.p,hb.data(),nsub*8,cudaMemcpyHostToDevice,c->st));
          k_suboff<<<(unsigned) ((nsub + 255) / 256),256,0,c->st>>>((u64 *) c->sub_off.p,(const u32 *) c->sub_par.p,
                                                                   (const u64 *) c->sub_base.p,(const u64 *) c->qoff.p,(long long) nsub); KCHECK();
          CompactParams cq = cp;
          cq.starts = (const u64 *) c->sub_s.p; cq.flags = (const u32 *) c->sub_f.p; cq.e_all = (const u32 *) c->sub_ea.p;
          cq.out_off = (const u64 *) c->sub_off.p; cq.nitems = (long long) nsub;
          k_compact_keys<NW><<<c->sms * 8,256,0,c->st>>>(cq,(PK *) c->pkeys.p,(uint16_t *) c->pcnts.p); KCHECK();
          CU(cudaStreamSynchronize(c->st));       /* hb is about to go out of scope */
        }
      rc = build_profile_lookup<(NW == 3) ? 2 : NW>(c,U);
      if (rc) return rc;
      c->ptab_n = (long long) U;
    }

  /* ---- table: scan the per-item pass counts, compact ------------------------------------------- */
  res->ntable = 0; res->table = NULL; res->table_dev = NULL;
  const int tw = c->kbytes + 2;
  if (direct)
    { u64 ntot_d;
      CU(cudaMemcpyAsync(&ntot_d,(const u64 *) c->off1.p + nb1,8,cudaMemcpyDeviceToHost,c->st));
      CU(cudaStreamSynchronize(c->st));
      res->ntable = (int64_t) ntot_d;
      res->table_dev = (const uint8_t *) c->table.p;
      if (fetch_table)
        { if (hout->ensure((size_t) ntot_d * tw + 64)) return set_err(FKGPU_E_NOMEM,"out of pinned host memory (table)");
          if (!d2h_chunked || hm.ovf_cnt > 0)       /* oversize items were finished after their range had been copied: take it all again */
            { if (c->d2h_pending) { CU(cudaStreamSynchronize(c->cst)); c->d2h_pending = false; }
              CU(cudaMemcpyAsync(hout->p,c->table.p,(size_t) ntot_d * tw,cudaMemcpyDeviceToHost,c->st));
            }
          res->table = (const uint8_t *) hout->p;
        }
    }
  else if (c->cfg.do_table > 0)
    { stage_begin(c,FKGPU_ST_COMPACT);
      int rc = run_large_scan<NW>(c,(const u32 *) c->epass.p,gmax,(u64 *) c->poff.p,&d_misc->total_pass);
      if (rc) return rc;
      CU(cudaMemcpyAsync(&hm,d_misc,sizeof(Misc),cudaMemcpyDeviceToHost,c->st));
      CU(cudaStreamSynchronize(c->st));
      if (c->table.ensure((size_t) hm.total_pass * tw + 64)) return set_err(FKGPU_E_NOMEM,"out of device memory (table of %llu entries)",hm.total_pass);
      CompactParams cp;
      cp.stage0 = Y; cp.stage1 = X; cp.stage_cnt = (const u32 *) c->scnt.p;
      cp.starts = gstart; cp.flags = NULL; cp.e_all = (const u32 *) c->eall.p; cp.out_off = (const u64 *) c->poff.p;
      cp.out = (uint8_t *) c->table.p; cp.nitems = gmax; cp.cutoff = (u32) c->cfg.do_table; cp.kbytes = c->kbytes;
      int grid = c->sms * 8;
      k_compact<NW><<<grid,256,0,c->st>>>(cp); KCHECK();
      if (!subs.empty())
        { size_t nsub = subs.size();
          CU(cudaMemcpyAsync(c->sub_base.p,sub_base_pass.data(),nsub*8,cudaMemcpyHostToDevice,c->st));
          k_suboff<<<(unsigned) ((nsub + 255) / 256),256,0,c->st>>>((u64 *) c->sub_off.p,(const u32 *) c->sub_par.p,
                                                                   (const u64 *) c->sub_base.p,(const u64 *) c->poff.p,(long long) nsub); KCHECK();
          CompactParams cq = cp;
          cq.starts = (const u64 *) c->sub_s.p; cq.flags = (const u32 *) c->sub_f.p; cq.e_all = (const u32 *) c->sub_ea.p;
          cq.out_off = (const u64 *) c->sub_off.p; cq.nitems = (long long) nsub;
          k_compact<NW><<<grid,256,0,c->st>>>(cq); KCHECK();
        }
      stage_end(c,FKGPU_ST_COMPACT);
      res->ntable = (int64_t) hm.total_pass;
      res->table_dev = (const uint8_t *) c->table.p;
      if (fetch_table)
        { if (hout->ensure((size_t) hm.total_pass * tw + 64)) return set_err(FKGPU_E_NOMEM,"out of pinned host memory (table)");
          CU(cudaMemcpyAsync(hout->p,c->table.p,(size_t) hm.total_pass * tw,cudaMemcpyDeviceToHost,c->st));
          res->table = (const uint8_t *) hout->p;
        }
    }

  /* ---- histogram + scalars --------------------------------------------------------------------- */
  u64 ntot;
  CU(cudaMemcpyAsync(c->h_hist,c->ghist.p,sizeof(c->h_hist),cudaMemcpyDeviceToHost,c->st));
  CU(cudaMemcpyAsync(&ntot,(const u64 *) c->off1.p + nb1,8,cudaMemcpyDeviceToHost,c->st));
  CU(cudaMemcpyAsync(&hm,d_misc,sizeof(Misc),cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));
  res->hist = c->h_hist;
  res->max_inst = (int64_t) hm.maxinst;
  res->ndistinct = (int64_t) hm.ndistinct;
  res->nkmers = (int64_t) ntot;
  c->last_ndist = (long long) hm.ndistinct;
  c->res_nw = (NW == 3) ? 2 : NW;
  return FKGPU_OK;
}

/*  the table's last key ranges may still be crossing PCIe on the copy stream: order the compute stream behind them, so
 *  that the end-of-count event (and the caller's synchronize) covers the whole table                                  */
static int d2h_join(fkgpu_ctx *c)
{ if (c->d2h_pending)
    { CU(cudaStreamWaitEvent(c->st,c->ev_d2h,0));
      c->d2h_pending = false;
    }
  return FKGPU_OK;
}

/*  result of a one-round count: the table is its own single run */
static void single_run(fkgpu_ctx *c, fkgpu_result *res)
{ c->run_n.assign(1,res->ntable);
  c->run_p.assign(1,res->table);
  res->nruns = 1; res->run_ntable = c->run_n.data(); res->run_table = c->run_p.data();
  c->st_rounds = 1;
}

static void make_kmask(int k, u32 *km)
{ for (int m = 0; m < 4; m++)
    { int lo = 32*m, hi = 32*(m+1);
      if (hi <= 2*k) km[m] = 0xffffffffu;
      else if (lo >= 2*k) km[m] = 0;
      else km[m] = 0xffffffffu << (hi - 2*k);
    }
}

static void choose_levels(long long nub, int *P1, int *P2)
{ unsigned long long want = (unsigned long long) std::max<long long>(1,nub / 256);
  int P = ilog2_ceil(want);
  if (P > 23) P = 23;
  int p1max = 11;
  { const char *e = getenv("FKGPU_P1"); if (e) p1max = std::max(1,std::min(11,atoi(e))); }
  *P1 = std::min(P,p1max);
  *P2 = std::min(13,P - *P1);            /* provisional; choose_p2 refines it from the level-1 histogram */
}

/*  Level-2 fan-out from the DENSEST level-1 bucket (a rank of the multi-GPU path, or skewed data, fills only part of
 *  the prefix space): fine buckets of that bucket average <= 512 records, well under the SC_CAP-SC_T+1 guarantee. */
static int choose_p2(fkgpu_ctx *c, int P1, int *P2)
{ const int nb1 = 1 << P1;
  std::vector<u64> h(nb1);
  CU(cudaMemcpyAsync(h.data(),c->hist1.p,(size_t) nb1 * 8,cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));
  u64 mx = 0;
  for (u64 x : h) mx = std::max(mx,x);
  int p2 = ilog2_ceil(std::max<u64>(1,(mx + 511) / 512));
  *P2 = std::min(13,p2);
  return FKGPU_OK;
}

static int prepare_small(fkgpu_ctx *c, int P1);
template<int NW> static int count_from_level1(fkgpu_ctx *c, long long nub, int P1, int P2, int fetch_table, fkgpu_result *res, void *bufX, void *bufY);

static int prepare_small(fkgpu_ctx *c, int P1)
{ const int nb1 = 1 << P1;
  if (c->hist1.ensure((size_t) (nb1 + 1) * 8) || c->off1.ensure((size_t) (nb1 + 1) * 8) || c->cur1.ensure((size_t) (nb1 + 1) * 8)
      || c->ghist.ensure(FKGPU_HIST_BINS * 8) || c->misc.ensure(sizeof(Misc)))
    return set_err(FKGPU_E_NOMEM,"out of device memory (histograms)");
  CU(cudaMemsetAsync(c->hist1.p,0,(size_t) (nb1 + 1) * 8,c->st));
  CU(cudaMemsetAsync(c->ghist.p,0,FKGPU_HIST_BINS * 8,c->st));
  CU(cudaMemsetAsync(c->misc.p,0,sizeof(Misc),c->st));
  return FKGPU_OK;
}

static int prepare_common(fkgpu_ctx *c, long long nub, int P1, bool needB = true, int rec_words = 0)
{ const size_t rb = (size_t) 8 * (rec_words ? rec_words : c->NW);
  if (c->bufA.ensure((size_t) (nub + 4) * rb) || (needB && c->bufB.ensure((size_t) (nub + 4) * rb)))
    return set_err(FKGPU_E_NOMEM,"out of device memory: two record buffers of %lld x %zu bytes",nub,rb);
  return prepare_small(c,P1);
}

/*  a multi-round count re-uses the stage events every round: bank what the finished round measured (stream idle) */
static void bank_times(fkgpu_ctx *c)
{ for (int s = 0; s < FKGPU_NSTAGES; s++)
    if (c->used[s])
      { float t = 0;
        if (cudaEventElapsedTime(&t,c->ev[2*s],c->ev[2*s+1]) == cudaSuccess) c->ms_bank[s] += t;
        else cudaGetLastError();
        c->used[s] = false;
      }
}

static void collect_times(fkgpu_ctx *c, fkgpu_result *res)
{ for (int s = 0; s < FKGPU_NSTAGES; s++)
    { c->ms[s] = c->ms_bank[s];
      c->ms_bank[s] = 0;
      if (c->used[s])
        { float t = 0;
          cudaEventElapsedTime(&t,c->ev[2*s],c->ev[2*s+1]);
          c->ms[s] += t;
        }
    }
  float tot = 0;
  cudaEventElapsedTime(&tot,c->ev[2*FKGPU_NSTAGES],c->ev[2*FKGPU_NSTAGES+1]);
  res->ms_pack = c->ms[FKGPU_ST_PACK];
  res->ms_total = tot;
  res->ms_count = tot - res->ms_pack;
}


/*  geometry of the persistent scan: the same (grid, tiles-per-CTA) must be used by HIST and SCATTER */
struct ScanGeom { long long ntiles, tpc; int grid; size_t smh, sms; };
static ScanGeom scan_geom(fkgpu_ctx *c, long long npos, int P1)
{ ScanGeom g;
  const int nb1 = 1 << P1;
  g.ntiles = (npos + SCAN_TILE - 1) / SCAN_TILE;
  long long want = (long long) c->sms * 4;
  g.grid = (int) std::max<long long>(1,std::min<long long>(g.ntiles,want));
  g.tpc = (g.ntiles + g.grid - 1) / std::max(1,g.grid);
  if (g.tpc < 1) g.tpc = 1;
  g.grid = (int) std::max<long long>(1,(g.ntiles + g.tpc - 1) / g.tpc);
  g.smh = (size_t) (SCAN_SEQW + SCAN_VALW + nb1 + (nb1 & 1)) * 4;
  g.sms = g.smh + (size_t) nb1 * 8;
  return g;
}

static void fill_scan_params(fkgpu_ctx *c, ScanParams &sp, const u32 *d_seq, const u32 *d_val, long long npos, int P1, const ScanGeom &g)
{ sp.seq = d_seq; sp.val = d_val; sp.npos = npos;
  sp.nvalw = (npos + 31) / 32; sp.nseqw = 2 * sp.nvalw;
  sp.k = c->cfg.kmer; sp.pbits = P1;
  make_kmask(c->cfg.kmer,sp.kmask);
  sp.ntiles = g.ntiles; sp.tpc = g.tpc;
  sp.cta_hist = (u32 *) c->ctah.p; sp.cta_off = (const u64 *) c->ctao.p; sp.off1 = NULL; sp.out = NULL;
}

/*  reads -> per-bucket totals in d_total[2^P1] (u64).  Leaves the per-CTA offsets in c->ctao for the scatter. */
template<int NW>
static int scan_hist(fkgpu_ctx *c, const u32 *d_seq, const u32 *d_val, long long npos, int P1, u64 *d_total)
{ const int nb1 = 1 << P1;
  ScanGeom g = scan_geom(c,npos,P1);
  if (c->ctah.ensure((size_t) g.grid * nb1 * 4) || c->ctao.ensure((size_t) g.grid * nb1 * 8))
    return set_err(FKGPU_E_NOMEM,"out of device memory (per-CTA histograms)");
  ScanParams sp;
  fill_scan_params(c,sp,d_seq,d_val,npos,P1,g);
  if (g.ntiles == 0)
    { CU(cudaMemsetAsync(d_total,0,(size_t) nb1 * 8,c->st));
      return FKGPU_OK;
    }
  CU(cudaFuncSetAttribute(k_scan<NW,false>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) g.smh));
  k_scan<NW,false><<<g.grid,SCAN_TPB,g.smh,c->st>>>(sp); KCHECK();
  k_colscan<<<(nb1 + 127) / 128,128,0,c->st>>>((const u32 *) c->ctah.p,(u64 *) c->ctao.p,d_total,g.grid,nb1); KCHECK();
  return FKGPU_OK;
}

template<int NW>
static int scan_scatter(fkgpu_ctx *c, const u32 *d_seq, const u32 *d_val, long long npos, int P1, const u64 *d_off1, void *d_out)
{ ScanGeom g = scan_geom(c,npos,P1);
  if (g.ntiles == 0) return FKGPU_OK;
  ScanParams sp;
  fill_scan_params(c,sp,d_seq,d_val,npos,P1,g);
  sp.off1 = d_off1; sp.out = d_out;
  const int nb1 = 1 << P1;
  static int variant = -1;
  if (variant < 0) { const char *e = getenv("FKGPU_SCAT"); variant = e ? atoi(e) : 1; }
  if (variant == 2)
    { CU(cudaFuncSetAttribute(k_scan<NW,true>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) g.sms));
      k_scan<NW,true><<<g.grid,SCAN_TPB,g.sms,c->st>>>(sp); KCHECK();
    }
  else
    { if (c->cur1.ensure((size_t) (nb1 + 1) * 8)) return set_err(FKGPU_E_NOMEM,"out of device memory (cursors)");
      CU(cudaMemcpyAsync(c->cur1.p,d_off1,(size_t) nb1 * 8,cudaMemcpyDeviceToDevice,c->st));
      sp.cursor = (u64 *) c->cur1.p;
      if (variant == 1)
        { CU(cudaFuncSetAttribute(k_scatter_tile<NW>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) g.sms));
          k_scatter_tile<NW><<<(unsigned) g.ntiles,SCAN_TPB,g.sms,c->st>>>(sp); KCHECK();
        }
      else
        { size_t sm = (size_t) (SCAN_SEQW + SCAN_VALW) * 4;
          k_scatter_atomic<NW><<<(unsigned) g.ntiles,SCAN_TPB,sm,c->st>>>(sp); KCHECK();
        }
    }
  return FKGPU_OK;
}

template<int NW>
static int count_packed_t(fkgpu_ctx *c, const u32 *d_seq, const u32 *d_val, long long npos, int fetch_table, fkgpu_result *res, bool own_total)
{ int P1, P2;
  choose_levels(npos,&P1,&P2);
  int rc = prepare_common(c,npos,P1);
  if (rc) return rc;
  const int nb1 = 1 << P1;

  if (own_total) cudaEventRecord(c->ev[2*FKGPU_NSTAGES],c->st);
  stage_begin(c,FKGPU_ST_SCANHIST);
  rc = scan_hist<NW>(c,d_seq,d_val,npos,P1,(u64 *) c->hist1.p);
  if (rc) return rc;
  stage_end(c,FKGPU_ST_SCANHIST);
  k_scan_small<<<1,1024,0,c->st>>>((const u64 *) c->hist1.p,(u64 *) c->off1.p,(u64 *) NULL,nb1); KCHECK();
  rc = choose_p2(c,P1,&P2);
  if (rc) return rc;
  stage_begin(c,FKGPU_ST_SCATTER);
  rc = scan_scatter<NW>(c,d_seq,d_val,npos,P1,(const u64 *) c->off1.p,c->bufA.p);
  if (rc) return rc;
  stage_end(c,FKGPU_ST_SCATTER);
  rc = count_from_level1<NW>(c,npos,P1,P2,fetch_table,res,c->bufA.p,c->bufB.p);
  if (rc) return rc;
  rc = d2h_join(c);
  if (rc) return rc;
  single_run(c,res);
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES+1],c->st);
  CU(cudaStreamSynchronize(c->st));
  collect_times(c,res);
  return FKGPU_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/*  super-mer path: reads -> 24-byte super-mer records bucketed by canonical minimizer -> per-bucket on-chip
 *  expansion + hash count -> (only if a table / profiles are wanted) key-order sort of the distinct entries.   */

static bool super_path_ok_k(int kmer)
{ static int forced = -1;
  if (forced < 0) { const char *e = getenv("FKGPU_PATH"); forced = e ? (strcmp(e,"records") == 0 ? 1 : (strcmp(e,"super") == 0 ? 2 : 0)) : 0; }
  if (forced == 1) return false;
  return kmer >= 18 && kmer <= FKGPU_MAX_K;
}
static bool super_path_ok(fkgpu_ctx *c) { return super_path_ok_k(c->cfg.kmer); }
/*  words of a distinct entry: (key | count in the low 16 bits) fits two words up to k = 56; beyond, the count takes a third */
static int entry_words(int kmer) { return kmer > 56 ? 3 : 2; }

struct SuperCounters { u64 nrec, nkmers, nent, spill_kmers, spill_cursor, sm_seen, sm_expanded; u32 fail, pad, nspill, pad2; };

/*  npos_total = positions over ALL ranks' read streams (multi-GPU: every rank must derive the same bucket-id width) */
static SuperGeom super_geom(int k, long long npos_total, long long npos_field = -1)
{ SuperGeom g;
  if (npos_field < 0) npos_field = npos_total;          /* positions the position field of a record must address */
  g.k = k; g.m = std::min(16,k - 8); g.w = k - g.m + 1;
  g.p2 = 1; while (2*g.p2 <= g.w) g.p2 <<= 1;
  const long long sest = std::max<long long>(1,npos_total / 10);        /* expected # of super-mers */
  int bbits = ilog2_ceil((unsigned long long) std::max<long long>(1,sest / 32));
  /* record = [bucket : bbits][# k-mers - 1 : 6][strand : 1][global position : pbits].  The bucket count grows with the input (22 bits up
     to 4 G positions, one more per doubling) so that buckets keep ~40 super-mers however many GPUs feed them.            */
  g.pbits = std::max(SUP_PBITS_MIN,ilog2_ceil((unsigned long long) npos_field + 1));
  int bcap = 22 + std::max(0,ilog2_ceil((unsigned long long) npos_total + 1) - 32);
  bcap = std::min(bcap,std::min(SUP_BBITS,64 - SUP_LBITS - 1 - g.pbits));
  if (bbits > bcap) bbits = bcap;
  { static int forced = -2;                 /* FKGPU_BBITS: force the bucket-id width (tests exercise the 8-GPU geometry on one GPU) */
    if (forced == -2) { const char *e = getenv("FKGPU_BBITS"); forced = e ? atoi(e) : -1; }
    if (forced >= 0) bbits = std::min(forced,std::min(SUP_BBITS,64 - SUP_LBITS - 1 - g.pbits));
  }
  static int sp1 = -1, sbb = -1;
  if (sp1 < 0) { const char *e = getenv("FKGPU_SP1"); sp1 = e ? atoi(e) : 11; const char *f = getenv("FKGPU_SBB"); sbb = f ? atoi(f) : 0; }
  if (sbb > 0 && bbits > sbb) bbits = sbb;
  g.bbits = bbits; g.P1 = std::min(bbits,sp1); g.P2 = bbits - g.P1;
  return g;
}

/*  stage A: reads -> super-mer records appended at out[0..) (count in d_cnt->nrec, k-mers covered in d_cnt->nkmers) */
static int super_scan_stage(fkgpu_ctx *c, const u32 *d_seq, const u32 *d_val, long long npos, const SuperGeom &g, u64 pos_offset,
                            u64 *out, u64 cap, SuperCounters *d_cnt)
{ const long long ntiles = (npos + SCAN_TILE - 1) / SCAN_TILE;
  if (ntiles <= 0) return FKGPU_OK;
  SuperParams sp;
  sp.seq = d_seq; sp.val = d_val; sp.npos = npos; sp.nvalw = (npos + 31) / 32; sp.nseqw = 2 * sp.nvalw;
  sp.k = g.k; sp.m = g.m; sp.w = g.w; sp.p2 = g.p2; sp.lmax = SUP_LMAX; sp.bbits = g.bbits ? g.bbits : 1;
  sp.out = out; sp.cap = cap; sp.counter = &d_cnt->nrec; sp.pos_offset = pos_offset; sp.pbits = g.pbits;
  const size_t sm = (size_t) (SCAN_SEQW + SCAN_VALW + 2*SUP_ROWS*SUP_RS) * 4;
#define SUPER_LAUNCH(P2V) do { \
    CU(cudaFuncSetAttribute(k_super<P2V>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) sm)); \
    k_super<P2V><<<(unsigned) ntiles,SCAN_TPB,sm,c->st>>>(sp); KCHECK(); } while (0)
  if (g.p2 == 8) SUPER_LAUNCH(8);
  else if (g.p2 == 16) SUPER_LAUNCH(16);
  else if (g.p2 == 32) SUPER_LAUNCH(32);
  else return set_err(FKGPU_E_UNSUPPORTED,"internal: minimizer window %d outside the super-mer kernel's range",g.w);
  return FKGPU_OK;
}

/*  level-1 histogram + partition of S super-mer records on the top b1 bucket-id bits: in -> out, starts in c->off1,
 *  per-bucket counts in c->hist1 (both 2^b1 (+1) u64)                                                                */
static int super_level1(fkgpu_ctx *c, const Key<1> *in, Key<1> *out, long long S, int b1)
{ const int n1 = 1 << b1;
  const size_t smh = (size_t) (n1 + (n1 & 1)) * 4, sms = smh + (size_t) n1 * 8;
  const long long nt = (S + TP_TILE(1) - 1) / TP_TILE(1);
  CU(cudaMemsetAsync(c->hist1.p,0,(size_t) (n1 + 1) * 8,c->st));
  if (nt > 0)
    { k_tilepart<1,false><<<(unsigned) nt,TP_TPB,smh,c->st>>>(in,NULL,(u64) S,b1,(u64 *) c->hist1.p); KCHECK(); }
  k_scan_small<<<1,1024,0,c->st>>>((const u64 *) c->hist1.p,(u64 *) c->off1.p,(u64 *) c->cur1.p,n1); KCHECK();
  if (nt > 0)
    { CU(cudaFuncSetAttribute(k_tilepart<1,true>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) sms));
      k_tilepart<1,true><<<(unsigned) nt,TP_TPB,sms,c->st>>>(in,out,(u64) S,b1,(u64 *) c->cur1.p); KCHECK();
    }
  return FKGPU_OK;
}

/*  Oversize bucket groups (listed by the bucket kernel in bp.spill_list): their nk k-mers are expanded to canonical records
 *  and counted by the record pipeline (tile partition -> MSD refine -> in-smem count, with its own refinement of skewed
 *  pieces and the saturating path for uniform ones) -- the histogram and the scalars accumulate in the same device
 *  counters, and the distinct k-mers at or above the cutoff rejoin the entries at ent[*nent ..) for the key-order sort.  */
template<int NW>
static int spill_count_t(fkgpu_ctx *c, const BucketParams &bp, const u32 *km, int kw, bool pay, u32 nspill, u64 nk,
                         void *ent, u64 ent_cap, SuperCounters *d_cnt, u64 *nent)
{ Misc *d_misc = (Misc *) c->misc.p;
  if (c->spillA.ensure((size_t) (nk + 4) * 8 * NW) || c->spillB.ensure((size_t) (nk + 4) * 8 * NW))
    return set_err(FKGPU_E_NOMEM,"out of device memory (%llu k-mers of oversize buckets)",nk);
  stage_begin(c,FKGPU_ST_SPILL);
  CU(cudaMemsetAsync(&d_cnt->spill_cursor,0,8,c->st));
#define SPILL_LAUNCH(KWV,PAYV) do { \
    k_spill_expand<NW,KWV,PAYV><<<nspill,256,0,c->st>>>(bp,km[KWV-1],(Key<NW> *) c->spillA.p,nk,&d_cnt->spill_cursor); KCHECK(); } while (0)
#define SPILL_LAUNCH_P(KWV) do { if (pay) SPILL_LAUNCH(KWV,true); else SPILL_LAUNCH(KWV,false); } while (0)
  if (kw == 2) SPILL_LAUNCH_P(2); else if (kw == 3) SPILL_LAUNCH_P(3); else SPILL_LAUNCH_P(4);
  int P1, P2;
  choose_levels((long long) nk,&P1,&P2);
  const int nb1 = 1 << P1;
  if (c->hist1.ensure((size_t) (nb1 + 1) * 8) || c->off1.ensure((size_t) (nb1 + 1) * 8) || c->cur1.ensure((size_t) (nb1 + 1) * 8))
    return set_err(FKGPU_E_NOMEM,"out of device memory");
  CU(cudaMemsetAsync(c->hist1.p,0,(size_t) (nb1 + 1) * 8,c->st));
  CU(cudaMemsetAsync(&d_misc->total_pass,0,16,c->st));                 /* total_pass, ovf_cnt, ticket: per-sort scratch */
  const size_t smh = (size_t) (nb1 + (nb1 & 1)) * 4, sms = smh + (size_t) nb1 * 8;
  const long long nt = ((long long) nk + TP_TILE(NW) - 1) / TP_TILE(NW);
  k_tilepart<NW,false><<<(unsigned) nt,TP_TPB,smh,c->st>>>((const Key<NW> *) c->spillA.p,NULL,nk,P1,(u64 *) c->hist1.p); KCHECK();
  k_scan_small<<<1,1024,0,c->st>>>((const u64 *) c->hist1.p,(u64 *) c->off1.p,(u64 *) c->cur1.p,nb1); KCHECK();
  int rc = choose_p2(c,P1,&P2);
  if (rc) return rc;
  CU(cudaFuncSetAttribute(k_tilepart<NW,true>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) sms));
  k_tilepart<NW,true><<<(unsigned) nt,TP_TPB,sms,c->st>>>((const Key<NW> *) c->spillA.p,(Key<NW> *) c->spillB.p,nk,P1,(u64 *) c->cur1.p); KCHECK();
  /* the record pipeline reads its options from the context: no profile table here, table cutoff = what the entries need */
  const fkgpu_config saved = c->cfg;
  const int wsaved = c->weighted;
  c->cfg.do_profile = 0; c->cfg.do_table = (ent != NULL) ? (int) bp.ent_min : 0; c->weighted = 0;
  fkgpu_result tmp; memset(&tmp,0,sizeof(tmp));
  rc = count_from_level1<NW>(c,(long long) nk,P1,P2,0,&tmp,c->spillB.p,c->spillA.p);
  c->cfg = saved; c->weighted = wsaved;
  if (rc) return rc;
  if (ent != NULL && tmp.ntable > 0)
    { if (*nent + (u64) tmp.ntable > ent_cap)
        return set_err(FKGPU_E_CUDA,"internal: distinct-entry buffer overflow after the oversize buckets (%llu + %lld > %llu)",*nent,(long long) tmp.ntable,ent_cap);
      const unsigned gr = (unsigned) ((tmp.ntable + 255) / 256);
      if (entry_words(c->cfg.kmer) == 3)
        k_table_to_entries<3><<<gr,256,0,c->st>>>((const uint8_t *) c->table.p,(u64) tmp.ntable,c->kbytes,(Key<3> *) ent,*nent,ent_cap);
      else
        k_table_to_entries<2><<<gr,256,0,c->st>>>((const uint8_t *) c->table.p,(u64) tmp.ntable,c->kbytes,(Key<2> *) ent,*nent,ent_cap);
      KCHECK();
      *nent += (u64) tmp.ntable;
    }
  stage_end(c,FKGPU_ST_SPILL);
  return FKGPU_OK;
}

/*  stage B: S super-mer records in `in` (consumed; `scratch` has room for S records too) -> bucket partition -> per-bucket
 *  on-chip expansion + hash count.  Histogram contributions go to c->ghist, scalars to d_misc / d_cnt, the distinct
 *  (key | count) entries to ent[0..ent_cap) when ent != NULL.  The records may point into the read streams of several
 *  ranks (seqr/pbase, nranks > 1): the bucket kernel then gathers the bases from peer memory over NVLink.
 *  super_bucket_range does this for the level-1 buckets [b0, b1) of records already partitioned on their top P1 bucket bits
 *  (l1recs, starts off1[0..n1]); Sr >= the number of records in the range.  A round of a multi-round count is one call.   */
static int super_bucket_range(fkgpu_ctx *c, const SuperGeom &g, const Key<1> *l1recs, Key<1> *l2recs, const u64 *off1, int b0, int b1, long long Sr,
                              const u32 *d_seq, int nranks, const u32 *const *seqr, const u64 *pbase, const void *payload, void *wait_event,
                              void *ent, u64 ent_cap, SuperCounters *d_cnt, SuperCounters *hc, Misc *hm, long long *ngroups)
/*  entries below the table cutoff never leave the chip unless profiles need every count */
{ Misc *d_misc = (Misc *) c->misc.p;
  const int bbits = g.bbits;
  const int p1 = bbits ? g.P1 : 0;
  const int nb = b1 - b0;
  const long long S = Sr;
  stage_begin(c,FKGPU_ST_SUPERREFINE);
  const u64 *offs = off1 + b0;
  const Key<1> *recs = l1recs;
  long long mbuckets = nb;
  if (bbits && g.P2 > 0)
    { mbuckets = (long long) nb << g.P2;
      if (c->off2.ensure((size_t) (mbuckets + 1) * 8)) return set_err(FKGPU_E_NOMEM,"out of device memory (off2)");
      const size_t sm = (size_t) REF_ST * REF_SLOT * sizeof(Key<1>) + (size_t) (1 << g.P2) * 4;
      CU(cudaFuncSetAttribute(k_refine<1>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) sm));
      CU(cudaMemsetAsync(&d_misc->ticket,0,4,c->st));
      k_refine<1><<<std::min(nb,c->sms * 2),REF_TPB,sm,c->st>>>(l1recs,l2recs,off1 + b0,nb,p1,g.P2,(u64 *) c->off2.p,&d_misc->ticket); KCHECK();
      offs = (const u64 *) c->off2.p; recs = l2recs;
    }
  stage_end(c,FKGPU_ST_SUPERREFINE);

  /* groups of whole buckets, ~TS super-mers each */
  static int bcvar = -1, tsv = 224;
  if (bcvar < 0)
    { const char *e = getenv("FKGPU_BC"); bcvar = (e && strcmp(e,"cta") == 0) ? 0 : 3;      /* FKGPU_BC=cta: the CTA-wide kernel without dedupe (A/B runs) */
      const char *f = getenv("FKGPU_TS"); if (f) tsv = std::max(8,atoi(f)); else if (bcvar == 3) tsv = 32;
    }
  const bool wide = entry_words(g.k) == 3;
  const int bcv = bcvar;
  const u32 TS = (u32) std::min(tsv,8192);
  const long long gmax = S / TS + 2;
  if (c->gstart.ensure((size_t) (gmax + 2) * 8)) return set_err(FKGPU_E_NOMEM,"out of device memory (groups)");
  u64 *gstart = (u64 *) c->gstart.p;
  k_fill_u64<<<(unsigned) ((gmax + 1 + 255) / 256),256,0,c->st>>>(gstart,gmax + 1,offs + mbuckets); KCHECK();
  k_groups<<<(unsigned) ((mbuckets + 1 + 255) / 256),256,0,c->st>>>(offs,mbuckets,TS,gstart,gmax); KCHECK();

  /* the payload may still be in flight (its all-to-all overlaps the partition above): order the counting kernel behind it */
  if (wait_event != NULL) CU(cudaStreamWaitEvent(c->st,(cudaEvent_t) wait_event,0));
  static int bigv = -1;
  if (bigv < 0) { const char *e = getenv("FKGPU_BIG"); bigv = e ? std::max(1,atoi(e)) : 0; }
  const u32 spill_cap = (u32) std::min<long long>(gmax,1 << 20);
  if (c->spill_list.ensure((size_t) spill_cap * 4)) return set_err(FKGPU_E_NOMEM,"out of device memory");
  BucketParams bp;
  u32 km[4]; make_kmask(g.k,km);
  const int kw = (2*g.k + 31) / 32;            /* 32-bit words of a key: 2 (k <= 32), 3 (k <= 48), 4 */
  stage_begin(c,FKGPU_ST_BUCKET);
  {
    bp.recs = (const u64 *) recs; bp.seq = d_seq; bp.starts = gstart; bp.ends = gstart + 1; bp.nitems = gmax; bp.k = g.k;
    bp.nranks = nranks; bp.pbits = g.pbits; bp.payload = (const uint4 *) payload;
    for (int r = 0; r < SUP_MAXRANKS; r++)
      { bp.seqr[r] = (nranks > 1 && r < nranks) ? seqr[r] : d_seq;
        bp.pbase[r] = (nranks > 1 && r < nranks) ? pbase[r] : 0;
      }
    bp.g_hist = (u64 *) c->ghist.p; bp.g_maxinst = &d_misc->maxinst; bp.g_ndistinct = &d_misc->ndistinct;
    bp.ent = ent; bp.ent_cap = ent_cap; bp.ent_counter = &d_cnt->nent;
    bp.ent_min = (u32) ((c->cfg.do_profile || c->cfg.do_table < 1) ? 1 : std::min(c->cfg.do_table,0x7fff));
    bp.g_fail = &d_cnt->fail;
    /* groups beyond this many super-mers leave the chip (FKGPU_BIG overrides; the old kernel streams everything) */
    bp.big = (u32) (bigv > 0 ? bigv : 2048);
    bp.g_stat = &d_cnt->sm_seen;
    bp.spill_cnt = &d_cnt->nspill; bp.spill_list = (u32 *) c->spill_list.p; bp.spill_cap = spill_cap; bp.spill_kmers = &d_cnt->spill_kmers;
    if (bcv == 3)
      { /* warp-private tables: a warp owns a group from load to emit */
#define BW_LAUNCH(KWV,PAYV,WIDEV) do { \
          CU(cudaFuncSetAttribute(k_bucket_count3<KWV,PAYV,WIDEV>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) BW_SMEM)); \
          int occ = 0; \
          CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,k_bucket_count3<KWV,PAYV,WIDEV>,BW_TPB,BW_SMEM)); \
          const long long grid = std::max<long long>(1,std::min<long long>((gmax + BW_WARPS - 1) / BW_WARPS,(long long) c->sms * std::max(1,occ))); \
          k_bucket_count3<KWV,PAYV,WIDEV><<<(unsigned) grid,BW_TPB,BW_SMEM,c->st>>>(bp,km[KWV-1]); KCHECK(); } while (0)
#define BW_LAUNCH_P(KWV,WIDEV) do { if (payload != NULL) BW_LAUNCH(KWV,true,WIDEV); else BW_LAUNCH(KWV,false,WIDEV); } while (0)
        if (kw == 2) BW_LAUNCH_P(2,false);
        else if (kw == 3) BW_LAUNCH_P(3,false);
        else if (wide) BW_LAUNCH_P(4,true);
        else BW_LAUNCH_P(4,false);
      }
    else if (bcv == 0)
      { /* persistent CTAs: one resident wave, groups dealt round-robin */
#define BK_LAUNCH(KWV,PAYV,WIDEV) do { \
          CU(cudaFuncSetAttribute(k_bucket_count2<KWV,PAYV,WIDEV>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) BK_SMEM)); \
          int occ = 0; \
          CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,k_bucket_count2<KWV,PAYV,WIDEV>,BK_TPB,BK_SMEM)); \
          const long long grid = std::max<long long>(1,std::min<long long>(gmax,(long long) c->sms * std::max(1,occ))); \
          k_bucket_count2<KWV,PAYV,WIDEV><<<(unsigned) grid,BK_TPB,BK_SMEM,c->st>>>(bp,km[KWV-1]); KCHECK(); } while (0)
#define BK_LAUNCH_P(KWV,WIDEV) do { if (payload != NULL) BK_LAUNCH(KWV,true,WIDEV); else BK_LAUNCH(KWV,false,WIDEV); } while (0)
        if (kw == 2) BK_LAUNCH_P(2,false);
        else if (kw == 3) BK_LAUNCH_P(3,false);
        else if (wide) BK_LAUNCH_P(4,true);
        else BK_LAUNCH_P(4,false);
      }
  }
  stage_end(c,FKGPU_ST_BUCKET);
  CU(cudaMemcpyAsync(hc,d_cnt,sizeof(*hc),cudaMemcpyDeviceToHost,c->st));
  CU(cudaMemcpyAsync(hm,d_misc,sizeof(*hm),cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));
  c->st_exp_acc += (long long) hc->sm_expanded;
  if (hc->fail) return set_err(FKGPU_E_CUDA,"internal: %u bucket groups could not be counted on chip",hc->fail);
  if (ent != NULL && hc->nent > ent_cap) return set_err(FKGPU_E_CUDA,"internal: distinct-entry buffer overflow (%llu > %llu)",hc->nent,ent_cap);
  if (hc->nspill > 0)
    { if (hc->nspill > spill_cap) return set_err(FKGPU_E_CUDA,"internal: %u oversize bucket groups exceed the list of %u",hc->nspill,spill_cap);
      const u32 nsp = hc->nspill; const u64 nkm = hc->spill_kmers;
      u64 ne = hc->nent;
      int rc = (g.k <= 32) ? spill_count_t<1>(c,bp,km,kw,payload != NULL,nsp,nkm,ent,ent_cap,d_cnt,&ne)
                           : spill_count_t<2>(c,bp,km,kw,payload != NULL,nsp,nkm,ent,ent_cap,d_cnt,&ne);
      if (rc) return rc;
      hc->nent = ne;
      CU(cudaMemcpyAsync(hm,d_misc,sizeof(*hm),cudaMemcpyDeviceToHost,c->st));
      CU(cudaStreamSynchronize(c->st));
      c->st_spill_acc += (long long) nkm;
    }
  *ngroups = gmax;
  return FKGPU_OK;
}


static int super_count_stage(fkgpu_ctx *c, const SuperGeom &g, Key<1> *in, Key<1> *scratch, long long S,
                             const u32 *d_seq, int nranks, const u32 *const *seqr, const u64 *pbase, const void *payload, void *wait_event,
                             void *ent, u64 ent_cap, SuperCounters *d_cnt, SuperCounters *hc, Misc *hm, long long *ngroups)
{ const int b1 = g.bbits ? g.P1 : 0;
  stage_begin(c,FKGPU_ST_SUPERPART);
  int rc = super_level1(c,in,scratch,S,b1);
  if (rc) return rc;
  stage_end(c,FKGPU_ST_SUPERPART);
  return super_bucket_range(c,g,scratch,in,(const u64 *) c->off1.p,0,1 << b1,S,d_seq,nranks,seqr,pbase,payload,wait_event,ent,ent_cap,d_cnt,hc,hm,ngroups);
}

/*  stage C: U distinct (key | count in the low 16 bits) entries in `ent` -> key order -> table.  The record pipeline in
 *  weighted mode; `ent` is consumed (second sort buffer), `other` holds U + 4 entries.                               */
template<int EW>
static int entries_sort_stage_t(fkgpu_ctx *c, void *ent, void *other, long long U, int fetch_table, fkgpu_result *res)
{ Misc *d_misc = (Misc *) c->misc.p;
  int q1, q2;
  choose_levels(U,&q1,&q2);
  const int n1 = 1 << q1;
  if (c->hist1.ensure((size_t) (n1 + 1) * 8) || c->off1.ensure((size_t) (n1 + 1) * 8) || c->cur1.ensure((size_t) (n1 + 1) * 8))
    return set_err(FKGPU_E_NOMEM,"out of device memory");
  CU(cudaMemsetAsync(c->hist1.p,0,(size_t) (n1 + 1) * 8,c->st));
  CU(cudaMemsetAsync(&d_misc->ticket,0,4,c->st));
  const size_t smh = (size_t) (n1 + (n1 & 1)) * 4, sms = smh + (size_t) n1 * 8;
  const long long nt = (U + TP_TILE(EW) - 1) / TP_TILE(EW);
  stage_begin(c,FKGPU_ST_ENTPART);
  if (nt > 0)
    { k_tilepart<EW,false><<<(unsigned) nt,TP_TPB,smh,c->st>>>((const Key<EW> *) ent,NULL,(u64) U,q1,(u64 *) c->hist1.p); KCHECK(); }
  k_scan_small<<<1,1024,0,c->st>>>((const u64 *) c->hist1.p,(u64 *) c->off1.p,(u64 *) c->cur1.p,n1); KCHECK();
  int rc = choose_p2(c,q1,&q2);
  if (rc) return rc;
  if (nt > 0)
    { CU(cudaFuncSetAttribute(k_tilepart<EW,true>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) sms));
      k_tilepart<EW,true><<<(unsigned) nt,TP_TPB,sms,c->st>>>((const Key<EW> *) ent,(Key<EW> *) other,(u64) U,q1,(u64 *) c->cur1.p); KCHECK();
    }
  stage_end(c,FKGPU_ST_ENTPART);
  c->weighted = 1;
  rc = count_from_level1<EW>(c,std::max<long long>(U,1),q1,q2,fetch_table,res,other,ent);
  c->weighted = 0;
  return rc;
}

static int entries_sort_stage(fkgpu_ctx *c, void *ent, void *other, long long U, int fetch_table, fkgpu_result *res)
{ return entry_words(c->cfg.kmer) == 3 ? entries_sort_stage_t<3>(c,ent,other,U,fetch_table,res)
                                       : entries_sort_stage_t<2>(c,ent,other,U,fetch_table,res);
}

/*  layout of the super-mer staging inside record buffer A (free until the final sort), for a buffer sized for nub positions */
struct SuperBufs { Key<1> *SA, *SB; u64 scap; };
static SuperBufs super_bufs(fkgpu_ctx *c, long long nub)
{ SuperBufs b;
  const size_t abytes = (size_t) (nub + 4) * 16;
  b.scap = (u64) (abytes / 2 / sizeof(u64)) - 8;
  b.SA = (Key<1> *) c->bufA.p;
  b.SB = (Key<1> *) ((char *) c->bufA.p + ((abytes / 2) & ~(size_t) 15));
  return b;
}

/*  prescanned: the streamed front end (stream_begin / stream_chunk) already sized the buffers for c->stream_cap positions,
 *  zeroed the counters and scanned every chunk into super-mer records during the ingest                                  */
static int count_packed_super(fkgpu_ctx *c, const u32 *d_seq, const u32 *d_val, long long npos, int fetch_table, fkgpu_result *res,
                              bool own_total, bool *fell_back, bool prescanned)
{ const long long nub = prescanned ? c->stream_nub : npos;
  const SuperGeom g = prescanned ? c->sgeom : super_geom(c->cfg.kmer,npos);
  *fell_back = false;
  int rc;
  if (!prescanned)
    { rc = prepare_common(c,nub,std::max(g.P1,1),true,entry_words(c->cfg.kmer));
      if (rc) return rc;
      if (c->segs.ensure(sizeof(SuperCounters))) return set_err(FKGPU_E_NOMEM,"out of device memory");
    }
  SuperCounters *d_cnt = (SuperCounters *) c->segs.p;
  if (!prescanned) CU(cudaMemsetAsync(d_cnt,0,sizeof(SuperCounters),c->st));
  const SuperBufs sb = super_bufs(c,nub);
  Key<1> *SA = sb.SA, *SB = sb.SB;
  const u64 scap = sb.scap;

  if (own_total) cudaEventRecord(c->ev[2*FKGPU_NSTAGES],c->st);
  if (!prescanned)
    { stage_begin(c,FKGPU_ST_SUPERSCAN);
      rc = super_scan_stage(c,d_seq,d_val,npos,g,0,(u64 *) SA,scap,d_cnt);
      if (rc) return rc;
      stage_end(c,FKGPU_ST_SUPERSCAN);
    }
  SuperCounters hc;
  CU(cudaMemcpyAsync(&hc,d_cnt,sizeof(hc),cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));
  if (hc.nrec > scap) { *fell_back = true; return FKGPU_OK; }        /* very short super-mers: use the record path */
  const long long S = (long long) hc.nrec;

  const bool want_entries = (c->cfg.do_table > 0) || c->cfg.do_profile;
  Misc hm;
  long long gmax = 0;
  /* the entry buffer was sized from nub positions; a streamed ingest may have delivered more than its reservation said
     (the stream is abandoned only beyond stream_cap): the distinct entries are bounded by the k-mers actually seen        */
  u64 ent_cap = (u64) nub;
  if (want_entries && hc.nkmers > ent_cap)
    { if (c->bufB.ensure((size_t) (hc.nkmers + 4) * 8 * entry_words(c->cfg.kmer)))
        return set_err(FKGPU_E_NOMEM,"out of device memory (entries of %llu k-mers)",hc.nkmers);
      ent_cap = hc.nkmers;
    }
  rc = super_count_stage(c,g,SA,SB,S,d_seq,1,NULL,NULL,NULL,NULL,want_entries ? c->bufB.p : NULL,ent_cap,d_cnt,&hc,&hm,&gmax);
  if (rc) return rc;
  static int verbose = -1;
  if (verbose < 0) { const char *e = getenv("FKGPU_VERBOSE"); verbose = e ? atoi(e) : 0; }
  if (verbose)
    fprintf(stderr,"[fkgpu] super-mer path: k=%d m=%d w=%d bbits=%d supermers=%llu (%.2f k-mers each) kmers=%llu distinct=%llu groups=%lld overflow classes=%u\n",
            g.k,g.m,g.w,g.bbits,hc.nrec,hc.nrec ? (double) hc.nkmers / hc.nrec : 0.,hc.nkmers,hm.ndistinct,gmax,hc.pad);

  res->ntable = 0; res->table = NULL; res->table_dev = NULL;
  c->ptab_n = 0;
  if (want_entries)
    { /* the super-mer records in buffer A are dead: it is the second sort buffer (and may have to grow with the entries) */
      if (c->bufA.ensure((size_t) (hc.nent + 4) * 8 * entry_words(c->cfg.kmer)))
        return set_err(FKGPU_E_NOMEM,"out of device memory (entry sort buffer of %llu entries)",hc.nent);
      rc = entries_sort_stage(c,c->bufB.p,c->bufA.p,(long long) hc.nent,fetch_table,res);
      if (rc) return rc;
      rc = d2h_join(c);
      if (rc) return rc;
    }
  else
    { CU(cudaMemcpyAsync(c->h_hist,c->ghist.p,sizeof(c->h_hist),cudaMemcpyDeviceToHost,c->st));
      CU(cudaStreamSynchronize(c->st));
      res->hist = c->h_hist;
    }
  res->max_inst = (int64_t) hm.maxinst;
  res->ndistinct = (int64_t) hm.ndistinct;
  res->nkmers = (int64_t) hc.nkmers;
  c->last_ndist = (long long) hm.ndistinct;
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES+1],c->st);
  CU(cudaStreamSynchronize(c->st));
  collect_times(c,res);
  c->last_path = 1;
  c->st_super = (long long) hc.nrec; c->st_ent = want_entries ? (long long) hc.nent : 0; c->st_groups = gmax;
  c->st_split = hc.pad; c->st_spill = c->st_spill_acc; c->st_expanded = c->st_exp_acc;
  single_run(c,res);
  return FKGPU_OK;
}

/*  Device memory the working buffers of a count may take: cfg.mem_limit (the host's -M), FKGPU_MEM_LIMIT (bytes; tests
 *  force multi-round counts of small inputs with it), else 0 = no limit beyond what is free.                          */
static size_t mem_limit_of(fkgpu_ctx *c)
{ const char *e = getenv("FKGPU_MEM_LIMIT");
  if (e && atoll(e) > 0) return (size_t) atoll(e);
  return c->cfg.mem_limit > 0 ? (size_t) c->cfg.mem_limit : 0;
}

/*  bytes the one-round super-mer count reserves for nub positions: two record buffers of entry width and the table */
static size_t one_round_bytes(fkgpu_ctx *c, long long nub)
{ const size_t ew = (size_t) 8 * entry_words(c->cfg.kmer);
  return (size_t) (nub + 4) * ew * 2 + (c->cfg.do_table > 0 ? (size_t) nub * (c->kbytes + 2) : 0) + ((size_t) 64 << 20);
}

/*  does the one-round working set fit?  (what the context already holds of it counts as available) */
static bool one_round_fits(fkgpu_ctx *c, long long nub)
{ const size_t need = one_round_bytes(c,nub), lim = mem_limit_of(c);
  if (lim && need > lim) return false;
  size_t fr = 0, tot = 0;
  if (cudaMemGetInfo(&fr,&tot) != cudaSuccess) { cudaGetLastError(); return true; }
  const size_t held = c->bufA.cap + c->bufB.cap + c->table.cap;
  return need <= (size_t) ((fr + held) * 0.94);
}

/*  Inputs beyond one round (the reference's NPARTS > 1: FastK.c:419-429 sizes the parts from -M, count.c:1337 loops over
 *  them, table.c:240-313 merges their sorted files).  The reads are scanned ONCE into super-mer records (0.7 B per k-mer)
 *  and partitioned on the top bucket bits; then contiguous ranges of level-1 buckets are counted one ROUND at a time, each
 *  sized so that even all-distinct k-mers fit the entry buffers (k-mers per level-1 bucket are summed exactly first).  A
 *  round's distinct entries are put in key order and leave as one sorted run, copied to pinned host memory while the next
 *  round is counted.  Runs hold disjoint key sets (a canonical k-mer lives in one minimizer bucket): no counts merge.    */
static int count_packed_super_rounds(fkgpu_ctx *c, const u32 *d_seq, const u32 *d_val, long long npos, int fetch_table, fkgpu_result *res,
                                     bool own_total)
{ const SuperGeom g = super_geom(c->cfg.kmer,npos);
  const int EW = entry_words(c->cfg.kmer), tw = c->kbytes + 2;
  const size_t EB = (size_t) 8 * EW;
  const bool want_entries = c->cfg.do_table > 0;
  if (c->cfg.do_profile)
    return set_err(FKGPU_E_NOMEM,"out of device memory: -p needs the whole table on the device and a multi-round count does not keep it");
  /* this path sizes the big buffers itself; what the context already holds of them is part of the budget (a second count of
     the same input re-uses them as they are: no cudaMalloc / cudaFree in the steady state)                                */
  size_t fr = 0, tot = 0;
  CU(cudaMemGetInfo(&fr,&tot));
  size_t budget = (size_t) ((fr + c->bufA.cap + c->bufB.cap + c->bufC.cap + c->table.cap) * 0.94);
  { const size_t lim = mem_limit_of(c); if (lim && lim < budget) budget = lim; }
  int rc = prepare_small(c,std::max(g.P1,1));
  if (rc) return rc;
  if (own_total) cudaEventRecord(c->ev[2*FKGPU_NSTAGES],c->st);

  /* ---- scan once: expected record density 2/(w+1) per position for random minimizers + 1/64 for the cuts, 50 % slack;
          k_super keeps counting past the capacity, so a denser input is rescanned once with the exact size            */
  SuperCounters hc;
  long long scap = (long long) (npos * std::min(1.0,(2.0 / (g.w + 1) + 1.0/64) * 1.5)) + (1 << 16);
  for (int attempt = 0; ; attempt++)
    { if ((size_t) (scap + 8) * 17 + budget / 8 > budget)
        return set_err(FKGPU_E_NOMEM,"out of device memory: %lld super-mer records do not fit the %zu MB budget",scap,budget >> 20);
      if (c->bufA.cap > (size_t) (scap + 8) * 40) c->bufA.release();          /* left over from a one-round count: far too large */
      if (c->bufA.ensure((size_t) (scap + 8) * 16) || c->segs.ensure(sizeof(SuperCounters)))
        return set_err(FKGPU_E_NOMEM,"out of device memory (super-mer records)");
      CU(cudaMemsetAsync(c->segs.p,0,sizeof(SuperCounters),c->st));
      stage_begin(c,FKGPU_ST_SUPERSCAN);
      rc = super_scan_stage(c,d_seq,d_val,npos,g,0,(u64 *) c->bufA.p,(u64) scap,(SuperCounters *) c->segs.p);
      if (rc) return rc;
      stage_end(c,FKGPU_ST_SUPERSCAN);
      CU(cudaMemcpyAsync(&hc,c->segs.p,sizeof(hc),cudaMemcpyDeviceToHost,c->st));
      CU(cudaStreamSynchronize(c->st));
      bank_times(c);
      if ((long long) hc.nrec <= scap) break;
      if (attempt) return set_err(FKGPU_E_CUDA,"internal: super-mer scan overflowed twice");
      scap = (long long) hc.nrec + 1024;
    }
  const long long S = (long long) hc.nrec;
  const u64 nkmers = hc.nkmers;
  Key<1> *SA = (Key<1> *) c->bufA.p;
  Key<1> *SB = (Key<1> *) ((char *) c->bufA.p + (((size_t) (scap + 8) * 8) & ~(size_t) 15));

  /* ---- level 1, then what every level-1 bucket holds: records (histogram) and k-mers (summed lengths) */
  const int b1 = g.bbits ? g.P1 : 0, n1 = 1 << b1;
  stage_begin(c,FKGPU_ST_SUPERPART);
  rc = super_level1(c,SA,SB,S,b1);
  if (rc) return rc;
  stage_end(c,FKGPU_ST_SUPERPART);
  if (c->roff1.ensure((size_t) (n1 + 1) * 8) || c->l1k.ensure((size_t) n1 * 8)) return set_err(FKGPU_E_NOMEM,"out of device memory");
  CU(cudaMemcpyAsync(c->roff1.p,c->off1.p,(size_t) (n1 + 1) * 8,cudaMemcpyDeviceToDevice,c->st));
  k_bucket_kmers<<<n1,256,0,c->st>>>((const u64 *) SB,(const u64 *) c->roff1.p,g.pbits,(u64 *) c->l1k.p); KCHECK();
  std::vector<u64> hrec(n1), hkm(n1);
  CU(cudaMemcpyAsync(hrec.data(),c->hist1.p,(size_t) n1 * 8,cudaMemcpyDeviceToHost,c->st));
  CU(cudaMemcpyAsync(hkm.data(),c->l1k.p,(size_t) n1 * 8,cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));
  bank_times(c);

  /* ---- entry buffers from what is left: ent + other + table per entry */
  long long ecap = 0;
  if (want_entries)
    { /* small stuff (bucket offsets, group lists, per-item counters) rides in the slack; 4 more bytes per entry for its count */
      const size_t used = c->bufA.cap + std::min((size_t) 256 << 20,budget / 8);
      { const size_t want = (size_t) ((budget > used ? budget - used : 0) / (2*EB + (size_t) tw + 4)) * EB;
        if (c->bufB.cap > want + want/2 + ((size_t) 64 << 20)) { c->bufB.release(); c->bufC.release(); c->table.release(); }   /* sized for another mode */
      }
      if (budget <= used) return set_err(FKGPU_E_NOMEM,"out of device memory: nothing left for the entry buffers (budget %zu MB)",budget >> 20);
      ecap = (long long) ((budget - used) / (2*EB + (size_t) tw + 4));
      if (ecap > (long long) nkmers) ecap = (long long) nkmers;
      u64 mx = 0;
      for (u64 x : hkm) mx = std::max(mx,x);
      if ((u64) ecap < mx)
        return set_err(FKGPU_E_NOMEM,"out of device memory: a level-1 bucket of %llu k-mers exceeds the %lld-entry buffers the %zu MB budget allows",
                       mx,ecap,budget >> 20);
      if (c->bufB.ensure((size_t) (ecap + 4) * EB) || c->bufC.ensure((size_t) (ecap + 4) * EB) || c->table.ensure((size_t) ecap * tw + 64))
        return set_err(FKGPU_E_NOMEM,"out of device memory (entry buffers of %lld entries)",ecap);
    }

  /* ---- rounds: greedy over the level-1 buckets, worst case (every k-mer distinct) fits the entry buffers */
  c->run_n.clear(); c->run_p.clear();
  res->ntable = 0; res->table = NULL; res->table_dev = NULL;
  c->ptab_n = 0;
  long long groups = 0, ents = 0; u32 splits = 0, fails = 0;
  Misc hm; memset(&hm,0,sizeof(hm));
  int nrounds = 0;
  for (int b0 = 0; b0 < n1; )
    { int bq = b0; u64 kr = 0, sr = 0;
      while (bq < n1 && (!want_entries || bq == b0 || kr + hkm[bq] <= (u64) ecap)) { kr += hkm[bq]; sr += hrec[bq]; bq++; }
      if (sr > 0)
        { if (c->segs.ensure(sizeof(SuperCounters))) return set_err(FKGPU_E_NOMEM,"out of device memory");
          SuperCounters *d_cnt = (SuperCounters *) c->segs.p;          /* the entry sort borrows this buffer: take it back */
          CU(cudaMemsetAsync(d_cnt,0,sizeof(SuperCounters),c->st));
          Misc *d_misc = (Misc *) c->misc.p;
          CU(cudaMemsetAsync(&d_misc->ovf_cnt,0,8,c->st));
          SuperCounters rc_h; long long gm = 0;
          rc = super_bucket_range(c,g,SB,SA,(const u64 *) c->roff1.p,b0,bq,(long long) sr,d_seq,1,NULL,NULL,NULL,NULL,
                                  want_entries ? c->bufB.p : NULL,(u64) ecap,d_cnt,&rc_h,&hm,&gm);
          if (rc) return rc;
          groups += gm; splits += rc_h.pad; fails += rc_h.fail;
          if (want_entries)
            { if ((size_t) nrounds >= c->h_runs.size()) c->h_runs.push_back(new PinBuf());
              c->h_out = c->h_runs[nrounds];
              fkgpu_result rr; memset(&rr,0,sizeof(rr));
              rc = entries_sort_stage(c,c->bufB.p,c->bufC.p,(long long) rc_h.nent,fetch_table,&rr);
              c->h_out = nullptr;
              if (rc) return rc;
              c->run_n.push_back(rr.ntable); c->run_p.push_back(rr.table);
              res->ntable += rr.ntable; ents += (long long) rc_h.nent;
            }
          bank_times(c);
          nrounds++;
        }
      b0 = bq;
    }
  if (nrounds == 0 && want_entries) { c->run_n.push_back(0); c->run_p.push_back(NULL); }
  rc = d2h_join(c);
  if (rc) return rc;
  CU(cudaMemcpyAsync(c->h_hist,c->ghist.p,sizeof(c->h_hist),cudaMemcpyDeviceToHost,c->st));
  CU(cudaMemcpyAsync(&hm,c->misc.p,sizeof(hm),cudaMemcpyDeviceToHost,c->st));
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES+1],c->st);
  CU(cudaStreamSynchronize(c->st));
  collect_times(c,res);
  static int verbose = -1;
  if (verbose < 0) { const char *e = getenv("FKGPU_VERBOSE"); verbose = e ? atoi(e) : 0; }
  if (verbose)
    fprintf(stderr,"[fkgpu] multi-round count: %d rounds over %d level-1 buckets, budget %zu MB, %lld super-mers, %llu k-mers, entry buffers %lld, %lld entries\n",
            nrounds,n1,budget >> 20,S,nkmers,ecap,ents);
  res->hist = c->h_hist;
  res->max_inst = (int64_t) hm.maxinst;
  res->ndistinct = (int64_t) hm.ndistinct;
  res->nkmers = (int64_t) nkmers;
  res->nruns = (int32_t) c->run_n.size();
  res->run_ntable = c->run_n.data(); res->run_table = c->run_p.data();
  if (res->nruns == 1) res->table = c->run_p[0];
  c->last_ndist = (long long) hm.ndistinct;
  c->last_path = 1;
  c->st_super = S; c->st_ent = ents; c->st_groups = groups; c->st_rounds = nrounds; c->st_split = splits; c->st_spill = c->st_spill_acc; c->st_expanded = c->st_exp_acc;
  (void) fails;
  return FKGPU_OK;
}

/*  Streamed front end.  stream_begin (first fkgpu_ingest after create / reset, c->mu held) sizes the device buffers from
 *  cfg.reserve_bases; stream_chunk (flush_tid, c->mu held, after c->st was made to wait for the chunk's copy) packs the
 *  chunk and, on the super-mer path, appends its super-mer records: chunks are self-contained (whole 0-terminated reads,
 *  a continued read re-delivers its k-1 overlap), so no k-mer spans two chunks.                                        */
static int stream_begin(fkgpu_ctx *c)
{ c->stream_on = false; c->stream_scan = false;
  static int off = -1;
  if (off < 0) { const char *e = getenv("FKGPU_NOSTREAM"); off = (e && atoi(e)) ? 1 : 0; }
  if (c->cfg.reserve_bases <= 0 || off) return FKGPU_OK;
  /* reserve_bases is only a hint (the reference-hosted shim extrapolates it from the first block): if the device
     cannot hold buffers of that size, forget the hint and let the buffers grow with what actually arrives          */
  const long long nub = (c->cfg.reserve_bases + c->cfg.reserve_bases/50 + (1 << 20)) & ~63ll;
  /* position space: a direct region ends when the next block does not fit, so up to one block per region is a zero gap */
  long long cap = (nub + c->cfg.reserve_bases/8 + (long long) c->tids.size() * (long long) c->chunk_bytes) & ~63ll;
  bool ok = (ascii_reserve(c,cap + 192) == 0);
  if (ok)
    { int64_t sw, vw;
      fkgpu_packed_words(cap,&sw,&vw);
      ok = !(c->seq.ensure((size_t) sw * 4) || c->val.ensure((size_t) vw * 4));
    }
  const bool scan = ok && !c->rel_table && c->mg == NULL && super_path_ok(c) && c->cfg.bc_prefix == 0 && (c->cfg.do_profile || one_round_fits(c,nub));
  if (scan)
    { c->sgeom = super_geom(c->cfg.kmer,cap);
      ok = (prepare_common(c,nub,std::max(c->sgeom.P1,1),true,entry_words(c->cfg.kmer)) == 0) && !c->segs.ensure(sizeof(SuperCounters));
    }
  if (!ok)
    { c->ascii.release(); c->seq.release(); c->val.release(); c->bufA.release(); c->bufB.release();
      c->cfg.reserve_bases = 0;
      g_err[0] = 0;
      return FKGPU_OK;
    }
  c->stream_cap = cap; c->stream_nub = nub;
  c->stream_on = true;
  if (scan)
    { CU(cudaMemsetAsync(c->segs.p,0,sizeof(SuperCounters),c->st));
      c->stream_scan = true;
    }
  return FKGPU_OK;
}

static int stream_chunk(fkgpu_ctx *c, long long off, long long len, long long scan_len)   /* multiples of 64 positions; scan_len <= len */
{ u32 *seq = (u32 *) c->seq.p + (off >> 4), *val = (u32 *) c->val.p + (off >> 5);
  const long long vw = len >> 5;
  if (vw <= 0) return FKGPU_OK;
  k_pack_ascii<<<(unsigned) ((vw + 255) / 256),256,0,c->st>>>((const uint4 *) ((const char *) c->ascii.p + off),len,seq,val,vw);
  KCHECK();
  if (c->stream_scan)
    { const SuperBufs sb = super_bufs(c,c->stream_nub);
      return super_scan_stage(c,seq,val,scan_len,c->sgeom,(u64) off,(u64 *) sb.SA,sb.scap,(SuperCounters *) c->segs.p);
    }
  return FKGPU_OK;
}

static int count_packed_any(fkgpu_ctx *c, const u32 *d_seq, const u32 *d_val, long long npos, int fetch_table, fkgpu_result *res, bool own_total,
                            bool prescanned = false)
{ c->last_path = 0;
  if (super_path_ok(c))
    { bool fb = false;
      if (!prescanned && !c->cfg.do_profile && !one_round_fits(c,npos))
        return count_packed_super_rounds(c,d_seq,d_val,npos,fetch_table,res,own_total);
      int rc = count_packed_super(c,d_seq,d_val,npos,fetch_table,res,own_total,&fb,prescanned);
      if (rc || !fb) return rc;
      memset(c->used,0,sizeof(c->used));
      own_total = true;
    }
  return (c->NW == 1) ? count_packed_t<1>(c,d_seq,d_val,npos,fetch_table,res,own_total)
                      : count_packed_t<2>(c,d_seq,d_val,npos,fetch_table,res,own_total);
}

static void init_result(fkgpu_ctx *c, fkgpu_result *res)
{ memset(res,0,sizeof(*res));
  memset(c->ms_bank,0,sizeof(c->ms_bank));
  c->st_spill_acc = 0; c->st_exp_acc = 0;
  res->kmer = c->cfg.kmer;
  res->kmer_bytes = c->kbytes;
  memset(c->used,0,sizeof(c->used));
}

extern "C" int fkgpu_count_packed(fkgpu_ctx *c, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos,
                                  int fetch_table, fkgpu_result *res)
{ if (c == NULL || res == NULL || (npos > 0 && (d_seq == NULL || d_val == NULL)))
    return set_err(FKGPU_E_ARG,"fkgpu_count_packed: NULL argument");
  if (npos < 0) return set_err(FKGPU_E_ARG,"fkgpu_count_packed: negative length");
  CU(cudaSetDevice(c->cfg.device));
  init_result(c,res);
  res->nbases = npos;
  return count_packed_any(c,d_seq,d_val,npos,fetch_table,res,true);
}

extern "C" int fkgpu_pack_ascii_dev(fkgpu_ctx *c, const char *d_ascii, int64_t npos, uint32_t *d_seq, uint32_t *d_val)
{ if (c == NULL || (npos > 0 && (d_ascii == NULL || d_seq == NULL || d_val == NULL)))
    return set_err(FKGPU_E_ARG,"fkgpu_pack_ascii_dev: NULL argument");
  CU(cudaSetDevice(c->cfg.device));
  if (((uintptr_t) d_ascii & 15) != 0) return set_err(FKGPU_E_ARG,"fkgpu_pack_ascii_dev: input must be 16-byte aligned");
  long long vw = (npos + 31) / 32;
  if (vw > 0)
    { k_pack_ascii<<<(unsigned) ((vw + 255) / 256),256,0,c->st>>>((const uint4 *) d_ascii,npos,d_seq,d_val,vw);
      KCHECK();
    }
  CU(cudaMemsetAsync(d_seq + 2*vw,0,FKGPU_PACK_PAD * 4,c->st));
  CU(cudaMemsetAsync(d_val + vw,0,FKGPU_PACK_PAD * 4,c->st));
  return FKGPU_OK;
}

static int count_packed_multi(fkgpu_ctx *c, const u32 *d_seq, const u32 *d_val, long long npos, int fetch_table, fkgpu_result *res);

extern "C" int fkgpu_finish(fkgpu_ctx *c, int fetch_table, fkgpu_result *res)
{ if (c == NULL || res == NULL) return set_err(FKGPU_E_ARG,"fkgpu_finish: NULL argument");
  if (c->finished) return set_err(FKGPU_E_STATE,"fkgpu_finish: already finished (use fkgpu_reset)");
  CU(cudaSetDevice(c->cfg.device));
  init_result(c,res);
  for (auto &t : c->tids)
    { int rc = flush_tid(c,t);
      if (rc) return rc;
      rc = close_region(c,t);
      if (rc) return rc;
    }
  CU(cudaStreamSynchronize(c->cst));
  c->finished = true;
  const long long npos = c->ascii_used;
  c->last_npos = npos;
  res->nbases = c->nbases;
  res->nreads = c->nreads;
  int64_t sw, vw;
  fkgpu_packed_words(npos,&sw,&vw);
  if (c->seq.ensure((size_t) sw * 4) || c->val.ensure((size_t) vw * 4))
    return set_err(FKGPU_E_NOMEM,"out of device memory (packed reads)");
  { std::lock_guard<std::mutex> lk(c->mu);
    if (ascii_reserve(c,npos + 64)) return set_err(FKGPU_E_NOMEM,"out of device memory (read buffer)");
  }
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES],c->st);
  stage_begin(c,FKGPU_ST_PACK);
  int rc = FKGPU_OK;
  if (c->stream_on)
    { /* every chunk was packed as it arrived; only the zero words behind the stream remain */
      const long long vwn = (npos + 31) / 32;
      CU(cudaMemsetAsync((u32 *) c->seq.p + 2*vwn,0,FKGPU_PACK_PAD * 4,c->st));
      CU(cudaMemsetAsync((u32 *) c->val.p + vwn,0,FKGPU_PACK_PAD * 4,c->st));
    }
  else
    rc = fkgpu_pack_ascii_dev(c,(const char *) c->ascii.p,npos,(uint32_t *) c->seq.p,(uint32_t *) c->val.p);
  if (rc) return rc;
  if (c->cfg.bc_prefix > 0 || c->cfg.do_profile)
    { /* read starts on the device, tid-major */
      std::vector<long long> rs;
      for (auto &t : c->tids)
        for (size_t i = 0; i < t.rstart.size(); i++)
          if (!t.rcont[i]) rs.push_back(t.rstart[i]);
      if (c->rstart_d.ensure(rs.size()*8 + 8)) return set_err(FKGPU_E_NOMEM,"out of device memory (read index)");
      if (!rs.empty())
        { CU(cudaMemcpyAsync(c->rstart_d.p,rs.data(),rs.size()*8,cudaMemcpyHostToDevice,c->st));
          CU(cudaStreamSynchronize(c->st));
          if (c->cfg.bc_prefix > 0)
            { k_mask_prefix<<<(unsigned) ((rs.size() + 255) / 256),256,0,c->st>>>((const long long *) c->rstart_d.p,(long long) rs.size(),
                                                                                  c->cfg.bc_prefix,npos,(u32 *) c->val.p);
              KCHECK();
            }
        }
    }
  stage_end(c,FKGPU_ST_PACK);
  if (c->rel_table)
    { /* -p:<table>: only profiles are produced, against the loaded table (FastK.c:328-337: no histogram, -t ignored) */
      cudaEventRecord(c->ev[2*FKGPU_NSTAGES+1],c->st);
      CU(cudaStreamSynchronize(c->st));
      collect_times(c,res);
      memset(c->h_hist,0,sizeof(c->h_hist));
      res->hist = c->h_hist;
      c->last_path = 0;
      single_run(c,res);
      return FKGPU_OK;
    }
  if (c->mg != NULL)
    return count_packed_multi(c,(const u32 *) c->seq.p,(const u32 *) c->val.p,npos,fetch_table,res);
  return count_packed_any(c,(const u32 *) c->seq.p,(const u32 *) c->val.p,npos,fetch_table,res,false,c->stream_on && c->stream_scan);
}

/* ------------------------------------------------------------------------------------------------ */
/*  multi-GPU stages                                                                                 */

extern "C" int fkgpu_prefix_hist(fkgpu_ctx *c, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos, int bits, uint64_t *d_hist)
{ if (c == NULL || d_hist == NULL || bits < 0 || bits > 11) return set_err(FKGPU_E_ARG,"fkgpu_prefix_hist: bad argument");
  CU(cudaSetDevice(c->cfg.device));
  int rc = (c->NW == 1) ? scan_hist<1>(c,d_seq,d_val,npos,bits,(u64 *) d_hist) : scan_hist<2>(c,d_seq,d_val,npos,bits,(u64 *) d_hist);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->st));
  return FKGPU_OK;
}

/*  must follow fkgpu_prefix_hist on the same (d_seq, d_val, npos, bits): it reuses the per-CTA regions computed there;
 *  d_hist = this rank's own histogram as returned by fkgpu_prefix_hist.                                              */
extern "C" int fkgpu_scatter_prefix(fkgpu_ctx *c, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos, int bits,
                                    const uint64_t *d_hist, void *d_records, int64_t cap_records, uint64_t *d_offsets)
{ if (c == NULL || d_hist == NULL || d_records == NULL || d_offsets == NULL || bits < 0 || bits > 11)
    return set_err(FKGPU_E_ARG,"fkgpu_scatter_prefix: bad argument");
  CU(cudaSetDevice(c->cfg.device));
  const int nb1 = 1 << bits;
  k_scan_small<<<1,1024,0,c->st>>>((const u64 *) d_hist,(u64 *) d_offsets,(u64 *) NULL,nb1); KCHECK();
  u64 tot;
  CU(cudaMemcpyAsync(&tot,d_offsets + nb1,8,cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));
  if ((int64_t) tot > cap_records) return set_err(FKGPU_E_ARG,"fkgpu_scatter_prefix: %llu records exceed the capacity %lld",tot,(long long) cap_records);
  int rc = (c->NW == 1) ? scan_scatter<1>(c,d_seq,d_val,npos,bits,(const u64 *) d_offsets,d_records)
                        : scan_scatter<2>(c,d_seq,d_val,npos,bits,(const u64 *) d_offsets,d_records);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->st));
  return FKGPU_OK;
}

template<int NW>
static int count_records_t(fkgpu_ctx *c, void *d_records, long long n, int fetch_table, fkgpu_result *res)
{ int P1, P2;
  choose_levels(n,&P1,&P2);
  int rc = prepare_common(c,n,P1,false);      /* the caller's buffer doubles as the second record buffer */
  if (rc) return rc;
  const int nb1 = 1 << P1;
  const size_t smh = (size_t) (nb1 + (nb1 & 1)) * 4;
  const size_t sms = smh + (size_t) nb1 * 8;
  const long long ntiles = (n + TP_TILE(NW) - 1) / TP_TILE(NW);
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES],c->st);
  if (ntiles > 0)
    { stage_begin(c,FKGPU_ST_L2HIST);
      k_tilepart<NW,false><<<(unsigned) ntiles,TP_TPB,smh,c->st>>>((const Key<NW> *) d_records,NULL,(u64) n,P1,(u64 *) c->hist1.p); KCHECK();
      stage_end(c,FKGPU_ST_L2HIST);
    }
  k_scan_small<<<1,1024,0,c->st>>>((const u64 *) c->hist1.p,(u64 *) c->off1.p,(u64 *) c->cur1.p,nb1); KCHECK();
  rc = choose_p2(c,P1,&P2);
  if (rc) return rc;
  if (ntiles > 0)
    { stage_begin(c,FKGPU_ST_SCATTER);
      CU(cudaFuncSetAttribute(k_tilepart<NW,true>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) sms));
      k_tilepart<NW,true><<<(unsigned) ntiles,TP_TPB,sms,c->st>>>((const Key<NW> *) d_records,(Key<NW> *) c->bufA.p,(u64) n,P1,(u64 *) c->cur1.p); KCHECK();
      stage_end(c,FKGPU_ST_SCATTER);
    }
  rc = count_from_level1<NW>(c,n,P1,P2,fetch_table,res,c->bufA.p,d_records);
  if (rc) return rc;
  rc = d2h_join(c);
  if (rc) return rc;
  single_run(c,res);
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES+1],c->st);
  CU(cudaStreamSynchronize(c->st));
  collect_times(c,res);
  return FKGPU_OK;
}

extern "C" int fkgpu_count_records(fkgpu_ctx *c, void *d_records, int64_t nrecords, int fetch_table, fkgpu_result *res)
{ if (c == NULL || res == NULL || nrecords < 0 || (nrecords > 0 && d_records == NULL))
    return set_err(FKGPU_E_ARG,"fkgpu_count_records: bad argument");
  CU(cudaSetDevice(c->cfg.device));
  init_result(c,res);
  return (c->NW == 1) ? count_records_t<1>(c,d_records,nrecords,fetch_table,res)
                      : count_records_t<2>(c,d_records,nrecords,fetch_table,res);
}

/* ------------------------------------------------------------------------------------------------ */
/*  multi-GPU stages of the super-mer path                                                           */

extern "C" int fkgpu_super_supported(int kmer)
{ return super_path_ok_k(kmer) ? 1 : 0; }

extern "C" int fkgpu_super_bucket_bits(int kmer, int64_t npos_total) { return super_geom(kmer,npos_total).bbits; }

extern "C" int fkgpu_reads_alloc(fkgpu_ctx *c, int64_t npos, uint32_t **d_seq, uint32_t **d_val)
{ if (c == NULL || d_seq == NULL || d_val == NULL || npos < 0) return set_err(FKGPU_E_ARG,"fkgpu_reads_alloc: bad argument");
  CU(cudaSetDevice(c->cfg.device));
  int64_t sw, vw;
  fkgpu_packed_words(npos,&sw,&vw);
  if (c->seq.ensure((size_t) sw * 4) || c->val.ensure((size_t) vw * 4))
    return set_err(FKGPU_E_NOMEM,"fkgpu_reads_alloc: out of device memory (packed reads of %lld positions)",(long long) npos);
  *d_seq = (uint32_t *) c->seq.p; *d_val = (uint32_t *) c->val.p;
  return FKGPU_OK;
}

extern "C" int fkgpu_ipc_export(fkgpu_ctx *c, const void *d_ptr, uint8_t *handle)
{ if (c == NULL || d_ptr == NULL || handle == NULL) return set_err(FKGPU_E_ARG,"fkgpu_ipc_export: NULL argument");
  CU(cudaSetDevice(c->cfg.device));
  static_assert(sizeof(cudaIpcMemHandle_t) == FKGPU_IPC_HANDLE_BYTES,"IPC handle size");
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h,(void *) d_ptr));
  memcpy(handle,&h,sizeof(h));
  return FKGPU_OK;
}

extern "C" int fkgpu_ipc_open(fkgpu_ctx *c, const uint8_t *handle, void **d_ptr)
{ if (c == NULL || d_ptr == NULL || handle == NULL) return set_err(FKGPU_E_ARG,"fkgpu_ipc_open: NULL argument");
  CU(cudaSetDevice(c->cfg.device));
  cudaIpcMemHandle_t h;
  memcpy(&h,handle,sizeof(h));
  CU(cudaIpcOpenMemHandle(d_ptr,h,cudaIpcMemLazyEnablePeerAccess));
  return FKGPU_OK;
}

extern "C" int fkgpu_ipc_close(fkgpu_ctx *c, void *d_ptr)
{ if (c == NULL || d_ptr == NULL) return set_err(FKGPU_E_ARG,"fkgpu_ipc_close: NULL argument");
  CU(cudaSetDevice(c->cfg.device));
  CU(cudaIpcCloseMemHandle(d_ptr));
  return FKGPU_OK;
}

extern "C" int fkgpu_super_scan(fkgpu_ctx *c, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos, int64_t npos_total,
                                int64_t pos_offset, const uint64_t **d_records, int64_t *nrecords, int64_t *nkmers,
                                const uint64_t **d_bucket_hist, const uint64_t **d_bucket_offsets, int32_t *hist_bits)
{ if (c == NULL || d_records == NULL || nrecords == NULL || nkmers == NULL || d_bucket_hist == NULL || d_bucket_offsets == NULL
      || hist_bits == NULL || npos < 0 || (npos > 0 && (d_seq == NULL || d_val == NULL)))
    return set_err(FKGPU_E_ARG,"fkgpu_super_scan: bad argument");
  if (!super_path_ok(c)) return set_err(FKGPU_E_UNSUPPORTED,"fkgpu_super_scan: k = %d is outside the super-mer path (18..64)",c->cfg.kmer);
  if (pos_offset + npos > npos_total || super_geom(c->cfg.kmer,npos_total).pbits > 64 - SUP_LBITS - 2)
    return set_err(FKGPU_E_ARG,"fkgpu_super_scan: positions [%lld,%lld) do not fit the declared total of %lld",(long long) pos_offset,
                   (long long) (pos_offset + npos),(long long) npos_total);
  CU(cudaSetDevice(c->cfg.device));
  memset(c->used,0,sizeof(c->used));
  const SuperGeom g = super_geom(c->cfg.kmer,npos_total);
  int rc = prepare_common(c,npos,std::max(g.P1,1),false,2);
  if (rc) return rc;
  if (c->segs.ensure(sizeof(SuperCounters))) return set_err(FKGPU_E_NOMEM,"out of device memory");
  SuperCounters *d_cnt = (SuperCounters *) c->segs.p;
  CU(cudaMemsetAsync(d_cnt,0,sizeof(SuperCounters),c->st));
  const size_t abytes = (size_t) (npos + 4) * 16;
  const u64 scap = (u64) (abytes / 2 / sizeof(u64)) - 8;
  Key<1> *SA = (Key<1> *) c->bufA.p;
  Key<1> *SB = (Key<1> *) ((char *) c->bufA.p + ((abytes / 2) & ~(size_t) 15));
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES],c->st);
  stage_begin(c,FKGPU_ST_SUPERSCAN);
  rc = super_scan_stage(c,d_seq,d_val,npos,g,(u64) pos_offset,(u64 *) SA,scap,d_cnt);
  if (rc) return rc;
  stage_end(c,FKGPU_ST_SUPERSCAN);
  SuperCounters hc;
  CU(cudaMemcpyAsync(&hc,d_cnt,sizeof(hc),cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));
  if (hc.nrec > scap) return set_err(FKGPU_E_UNSUPPORTED,"fkgpu_super_scan: %llu super-mers exceed the staging buffer (very short super-mers)",hc.nrec);
  const int b1 = g.bbits ? g.P1 : 0;
  stage_begin(c,FKGPU_ST_SUPERPART);
  rc = super_level1(c,SA,SB,(long long) hc.nrec,b1);
  if (rc) return rc;
  stage_end(c,FKGPU_ST_SUPERPART);
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES+1],c->st);
  CU(cudaStreamSynchronize(c->st));
  fkgpu_result tmp; collect_times(c,&tmp);
  *d_records = (const uint64_t *) SB; *nrecords = (int64_t) hc.nrec; *nkmers = (int64_t) hc.nkmers;
  *d_bucket_hist = (const uint64_t *) c->hist1.p; *d_bucket_offsets = (const uint64_t *) c->off1.p; *hist_bits = b1;
  return FKGPU_OK;
}

extern "C" int fkgpu_super_payload(fkgpu_ctx *c, const uint32_t *d_seq, int64_t pos_offset, int64_t npos_total,
                                   const uint64_t *d_records, int64_t nrecords, void *d_payload)
{ if (c == NULL || nrecords < 0 || (nrecords > 0 && (d_seq == NULL || d_records == NULL || d_payload == NULL)))
    return set_err(FKGPU_E_ARG,"fkgpu_super_payload: bad argument");
  if (!super_path_ok(c)) return set_err(FKGPU_E_UNSUPPORTED,"fkgpu_super_payload: k = %d is outside the super-mer path (18..64)",c->cfg.kmer);
  CU(cudaSetDevice(c->cfg.device));
  const SuperGeom g = super_geom(c->cfg.kmer,npos_total);
  if (nrecords > 0)
    { k_materialise<<<(unsigned) ((nrecords + 255) / 256),256,0,c->st>>>((const u64 *) d_records,(long long) nrecords,g.pbits,(u64) pos_offset,
                                                                        g.k,d_seq,(uint4 *) d_payload); KCHECK();
    }
  CU(cudaStreamSynchronize(c->st));
  return FKGPU_OK;
}

extern "C" int fkgpu_super_count(fkgpu_ctx *c, uint64_t *d_records, int64_t nrecords, int64_t npos_total, int32_t nranks,
                                 const uint32_t *const *seq_of_rank, const int64_t *pos_base, const void *d_payload,
                                 void *payload_ready_event, int want_entries, fkgpu_result *res, const void **d_entries, int64_t *nentries)
{ if (c == NULL || res == NULL || nrecords < 0 || (nrecords > 0 && d_records == NULL)
      || (d_payload == NULL && (nranks < 1 || nranks > SUP_MAXRANKS || seq_of_rank == NULL || pos_base == NULL))
      || (want_entries && (d_entries == NULL || nentries == NULL)))
    return set_err(FKGPU_E_ARG,"fkgpu_super_count: bad argument");
  if (!super_path_ok(c)) return set_err(FKGPU_E_UNSUPPORTED,"fkgpu_super_count: k = %d is outside the super-mer path (18..64)",c->cfg.kmer);
  CU(cudaSetDevice(c->cfg.device));
  init_result(c,res);
  const SuperGeom g = super_geom(c->cfg.kmer,npos_total);
  int rc = prepare_small(c,std::max(g.P1,1));
  if (rc) return rc;
  if (c->bufA.ensure((size_t) (nrecords + 8) * 8) || c->segs.ensure(sizeof(SuperCounters)) || c->bsum.ensure(8))
    return set_err(FKGPU_E_NOMEM,"out of device memory (super-mer scratch of %lld records)",(long long) nrecords);
  SuperCounters *d_cnt = (SuperCounters *) c->segs.p;
  CU(cudaMemsetAsync(d_cnt,0,sizeof(SuperCounters),c->st));
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES],c->st);
  /* k-mers covered by the received records = capacity bound of the distinct-entry buffer */
  u64 nk = 0;
  CU(cudaMemsetAsync(c->bsum.p,0,8,c->st));
  if (nrecords > 0)
    { k_sum_lengths<<<c->sms * 4,256,0,c->st>>>((const u64 *) d_records,(long long) nrecords,super_geom(c->cfg.kmer,npos_total).pbits,(u64 *) c->bsum.p); KCHECK(); }
  CU(cudaMemcpyAsync(&nk,c->bsum.p,8,cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));
  void *ent = NULL;
  if (want_entries)
    { if (c->bufB.ensure((size_t) (nk + 4) * 8 * entry_words(c->cfg.kmer))) return set_err(FKGPU_E_NOMEM,"out of device memory (entries of %llu k-mers)",nk);
      ent = c->bufB.p;
    }
  const u32 *seqr[SUP_MAXRANKS]; u64 pb[SUP_MAXRANKS];
  for (int r = 0; r < SUP_MAXRANKS; r++)
    { seqr[r] = d_payload ? NULL : seq_of_rank[r < nranks ? r : 0]; pb[r] = d_payload ? 0 : (u64) pos_base[r < nranks ? r : 0]; }
  if (d_payload != NULL && ((unsigned long long) nrecords * 8ull) >> g.pbits)
    return set_err(FKGPU_E_ARG,"fkgpu_super_count: %lld payload strings exceed the position field of a record",(long long) nrecords);
  if (d_payload != NULL && nrecords > 0)
    { k_reindex<<<(unsigned) ((nrecords + 255) / 256),256,0,c->st>>>((u64 *) d_records,(long long) nrecords,g.pbits); KCHECK(); }
  SuperCounters hc; Misc hm; long long gmax = 0;
  rc = super_count_stage(c,g,(Key<1> *) d_records,(Key<1> *) c->bufA.p,(long long) nrecords,seqr[0],d_payload ? 1 : nranks,seqr,pb,d_payload,
                         d_payload ? payload_ready_event : NULL,ent,nk,d_cnt,&hc,&hm,&gmax);
  if (rc) return rc;
  CU(cudaMemcpyAsync(c->h_hist,c->ghist.p,sizeof(c->h_hist),cudaMemcpyDeviceToHost,c->st));
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES+1],c->st);
  CU(cudaStreamSynchronize(c->st));
  collect_times(c,res);
  { const char *e = getenv("FKGPU_VERBOSE");
    if (e && atoi(e))
      fprintf(stderr,"[fkgpu] super_count: records=%lld kmers=%llu distinct=%llu entries=%llu groups=%lld overflow classes=%u bucket %.2f ms\n",
              (long long) nrecords,nk,hm.ndistinct,hc.nent,gmax,hc.pad,c->ms[FKGPU_ST_BUCKET]);
  }
  res->hist = c->h_hist;
  res->max_inst = (int64_t) hm.maxinst;
  res->ndistinct = (int64_t) hm.ndistinct;
  res->nkmers = (int64_t) nk;
  c->last_path = 1; c->st_super = (long long) nrecords; c->st_ent = want_entries ? (long long) hc.nent : 0; c->st_groups = gmax;
  if (want_entries) { *d_entries = ent; *nentries = (int64_t) hc.nent; }
  return FKGPU_OK;
}

template<int EW>
static int entries_partition_t(fkgpu_ctx *c, const void *d_entries, int64_t n, int bits, void *d_out, uint64_t *d_hist, uint64_t *d_offsets)
{ const int n1 = 1 << bits;
  if (c->cur1.ensure((size_t) (n1 + 1) * 8)) return set_err(FKGPU_E_NOMEM,"out of device memory (cursors)");
  const size_t smh = (size_t) (n1 + (n1 & 1)) * 4, sms = smh + (size_t) n1 * 8;
  const long long nt = (n + TP_TILE(EW) - 1) / TP_TILE(EW);
  CU(cudaMemsetAsync(d_hist,0,(size_t) n1 * 8,c->st));
  if (nt > 0)
    { k_tilepart<EW,false><<<(unsigned) nt,TP_TPB,smh,c->st>>>((const Key<EW> *) d_entries,NULL,(u64) n,bits,(u64 *) d_hist); KCHECK(); }
  k_scan_small<<<1,1024,0,c->st>>>((const u64 *) d_hist,(u64 *) d_offsets,(u64 *) c->cur1.p,n1); KCHECK();
  if (nt > 0)
    { CU(cudaFuncSetAttribute(k_tilepart<EW,true>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int) sms));
      k_tilepart<EW,true><<<(unsigned) nt,TP_TPB,sms,c->st>>>((const Key<EW> *) d_entries,(Key<EW> *) d_out,(u64) n,bits,(u64 *) c->cur1.p); KCHECK();
    }
  CU(cudaStreamSynchronize(c->st));
  return FKGPU_OK;
}

extern "C" int fkgpu_entry_bytes(int kmer) { return 8 * entry_words(kmer); }

extern "C" int fkgpu_entries_partition(fkgpu_ctx *c, const void *d_entries, int64_t n, int bits, void *d_out,
                                       uint64_t *d_hist, uint64_t *d_offsets)
{ if (c == NULL || n < 0 || (n > 0 && (d_entries == NULL || d_out == NULL)) || d_hist == NULL || d_offsets == NULL || bits < 0 || bits > 11)
    return set_err(FKGPU_E_ARG,"fkgpu_entries_partition: bad argument");
  CU(cudaSetDevice(c->cfg.device));
  return entry_words(c->cfg.kmer) == 3 ? entries_partition_t<3>(c,d_entries,n,bits,d_out,d_hist,d_offsets)
                                       : entries_partition_t<2>(c,d_entries,n,bits,d_out,d_hist,d_offsets);
}

extern "C" int fkgpu_entries_sort(fkgpu_ctx *c, void *d_entries, int64_t n, int fetch_table, fkgpu_result *res)
{ if (c == NULL || res == NULL || n < 0 || (n > 0 && d_entries == NULL)) return set_err(FKGPU_E_ARG,"fkgpu_entries_sort: bad argument");
  CU(cudaSetDevice(c->cfg.device));
  init_result(c,res);
  int rc = prepare_small(c,1);
  if (rc) return rc;
  if (c->bufA.ensure((size_t) (n + 4) * 8 * entry_words(c->cfg.kmer))) return set_err(FKGPU_E_NOMEM,"out of device memory (entry sort buffer)");
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES],c->st);
  rc = entries_sort_stage(c,d_entries,c->bufA.p,(long long) n,fetch_table,res);
  if (rc) return rc;
  rc = d2h_join(c);
  if (rc) return rc;
  single_run(c,res);
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES+1],c->st);
  CU(cudaStreamSynchronize(c->st));
  collect_times(c,res);
  res->nkmers = 0; res->ndistinct = n;
  return FKGPU_OK;
}


/* ------------------------------------------------------------------------------------------------ */
/*  multi-GPU count inside the library (fkgpu_multi.cuh has the design)                              */

static void mg_destroy(fkgpu_ctx *c)
{ MultiState *m = c->mg;
  if (m == NULL) return;
  fkmg::Api *na = fkmg::api();
  if (m->comm && na) na->CommDestroy(m->comm);
  if (m->ev_rec) cudaEventDestroy(m->ev_rec);
  if (m->ev_pay) cudaEventDestroy(m->ev_pay);
  DevBuf *bufs[] = { &m->small,&m->payload,&m->rrec,&m->rscr,&m->rpay,&m->epart,&m->erecv };
  for (auto b : bufs) b->release();
  delete m;
  c->mg = NULL;
}

extern "C" int fkgpu_comm_id(uint8_t *id)
{ fkmg::Api *na = fkmg::api();
  if (id == NULL) return set_err(FKGPU_E_ARG,"fkgpu_comm_id: NULL argument");
  if (na == NULL) return set_err(FKGPU_E_UNSUPPORTED,"fkgpu_comm_id: libnccl.so.2 could not be loaded (set FKGPU_NCCL_LIB)");
  fkmg::UniqueId u;
  NC(na->GetUniqueId(&u));
  static_assert(sizeof(u) == FKGPU_COMM_ID_BYTES,"NCCL unique id size");
  memcpy(id,&u,sizeof(u));
  return FKGPU_OK;
}

extern "C" int fkgpu_comm_init(fkgpu_ctx *c, int nranks, int rank, const uint8_t *id)
{ if (c == NULL || id == NULL || nranks < 1 || nranks > 2*SUP_MAXRANKS || rank < 0 || rank >= nranks) return set_err(FKGPU_E_ARG,"fkgpu_comm_init: bad argument (at most %d ranks)",2*SUP_MAXRANKS);
  fkmg::Api *na = fkmg::api();
  if (na == NULL) return set_err(FKGPU_E_UNSUPPORTED,"fkgpu_comm_init: libnccl.so.2 could not be loaded (set FKGPU_NCCL_LIB)");
  CU(cudaSetDevice(c->cfg.device));
  if (c->mg) mg_destroy(c);
  MultiState *m = new (std::nothrow) MultiState();
  if (m == NULL) return set_err(FKGPU_E_NOMEM,"fkgpu_comm_init: out of host memory");
  fkmg::UniqueId u;
  memcpy(&u,id,sizeof(u));
  int r = na->CommInitRank(&m->comm,nranks,u,rank);
  if (r != 0) { delete m; return set_err(FKGPU_E_CUDA,"ncclCommInitRank failed: %s",na->GetErrorString ? na->GetErrorString(r) : "NCCL error"); }
  m->nranks = nranks; m->rank = rank;
  CU(cudaEventCreateWithFlags(&m->ev_rec,cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&m->ev_pay,cudaEventDisableTiming));
  c->mg = m;
  return FKGPU_OK;
}

extern "C" int fkgpu_comm_info(fkgpu_ctx *c, int64_t *v, int64_t *table_sizes)
{ if (c == NULL || v == NULL) return set_err(FKGPU_E_ARG,"fkgpu_comm_info: NULL argument");
  if (c->mg == NULL) return set_err(FKGPU_E_STATE,"fkgpu_comm_info: no communicator (fkgpu_comm_init)");
  MultiState *m = c->mg;
  int64_t off = 0;
  for (int r = 0; r < m->rank && r < (int) m->table_sizes.size(); r++) off += m->table_sizes[r];
  v[0] = m->nranks; v[1] = m->rank; v[2] = m->global_ntable; v[3] = off; v[4] = m->sent_records; v[5] = m->sent_entries;
  if (table_sizes)
    for (int r = 0; r < m->nranks; r++) table_sizes[r] = r < (int) m->table_sizes.size() ? m->table_sizes[r] : 0;
  return FKGPU_OK;
}

/*  every rank contributes n u64 values; -> all of them, rank-major, on the host (one all-gather, one synchronize) */
static int mg_gather_u64(fkgpu_ctx *c, const u64 *mine, int n, std::vector<u64> &all)
{ MultiState *m = c->mg;
  fkmg::Api *na = fkmg::api();
  const int W = m->nranks;
  if (m->small.ensure((size_t) (W + 1) * n * 8 + 64)) return set_err(FKGPU_E_NOMEM,"out of device memory");
  u64 *d_mine = (u64 *) m->small.p, *d_all = d_mine + n;
  CU(cudaMemcpyAsync(d_mine,mine,(size_t) n * 8,cudaMemcpyHostToDevice,c->st));
  NC(na->AllGather(d_mine,d_all,(size_t) n,fkmg::kUint64,m->comm,c->st));
  all.resize((size_t) W * n);
  CU(cudaMemcpyAsync(all.data(),d_all,(size_t) W * n * 8,cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));
  return FKGPU_OK;
}

/*  variable-size all-to-all of elements of eb bytes: slice r of `send` (element offset soff[r], scnt[r] elements) goes to rank
    r, the slice of rank r lands at element offset roff[r] of `recv`; one NCCL group, the own slice is a device copy          */
static int mg_alltoall(fkgpu_ctx *c, const void *send, const std::vector<u64> &soff, const std::vector<u64> &scnt,
                       void *recv, const std::vector<u64> &roff, const std::vector<u64> &rcnt, size_t eb, cudaStream_t st)
{ MultiState *m = c->mg;
  fkmg::Api *na = fkmg::api();
  NC(na->GroupStart());
  for (int r = 0; r < m->nranks; r++)
    { if (r == m->rank) continue;
      if (scnt[r]) NC(na->Send((const char *) send + soff[r] * eb,(size_t) scnt[r] * eb,fkmg::kUint8,r,m->comm,st));
      if (rcnt[r]) NC(na->Recv((char *) recv + roff[r] * eb,(size_t) rcnt[r] * eb,fkmg::kUint8,r,m->comm,st));
    }
  NC(na->GroupEnd());
  if (scnt[m->rank])
    CU(cudaMemcpyAsync((char *) recv + roff[m->rank] * eb,(const char *) send + soff[m->rank] * eb,(size_t) scnt[m->rank] * eb,
                       cudaMemcpyDeviceToDevice,st));
  return FKGPU_OK;
}

/*  plan of one exchange: what this rank sends to everyone (from the group starts `off` of its locally partitioned items and
    the agreed cut points `beg`), what it receives (all-gather of the send counts)                                          */
struct MgPlan { std::vector<u64> soff, scnt, roff, rcnt; u64 nsend = 0, nrecv = 0; };

static int mg_plan(fkgpu_ctx *c, const std::vector<u64> &off, const std::vector<int> &beg, MgPlan &pl)
{ MultiState *m = c->mg;
  const int W = m->nranks;
  pl.soff.resize(W); pl.scnt.resize(W); pl.roff.resize(W); pl.rcnt.resize(W);
  for (int r = 0; r < W; r++) { pl.soff[r] = off[beg[r]]; pl.scnt[r] = off[beg[r+1]] - off[beg[r]]; }
  std::vector<u64> all;
  int rc = mg_gather_u64(c,pl.scnt.data(),W,all);
  if (rc) return rc;
  pl.nsend = 0; pl.nrecv = 0;
  for (int r = 0; r < W; r++)
    { pl.rcnt[r] = all[(size_t) r * W + m->rank];
      pl.roff[r] = pl.nrecv; pl.nrecv += pl.rcnt[r];
      pl.nsend += pl.scnt[r];
    }
  return FKGPU_OK;
}

static int count_packed_multi(fkgpu_ctx *c, const u32 *d_seq, const u32 *d_val, long long npos, int fetch_table, fkgpu_result *res)
{ MultiState *m = c->mg;
  fkmg::Api *na = fkmg::api();
  const int W = m->nranks, me = m->rank;
  if (!super_path_ok(c)) return set_err(FKGPU_E_UNSUPPORTED,"multi-GPU count: k = %d is outside the super-mer path (18..64)",c->cfg.kmer);
  if (c->cfg.do_profile) return set_err(FKGPU_E_UNSUPPORTED,"multi-GPU count: -p needs the whole table on one device");
  const int EW = entry_words(c->cfg.kmer);
  const size_t EB = (size_t) 8 * EW;
  const bool want_entries = c->cfg.do_table > 0;
  memset(c->used,0,sizeof(c->used));
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES],c->st);
  /* FKGPU_MG_TIMING=1: host wall clock of every phase, both streams drained at each mark (diagnostic: it serialises the overlap) */
  static int tl_on = -1;
  if (tl_on < 0) { const char *e = getenv("FKGPU_MG_TIMING"); tl_on = (e && atoi(e)) ? 1 : 0; }
  struct timespec tl_t0; clock_gettime(CLOCK_MONOTONIC,&tl_t0);
  char tl_buf[1024]; int tl_len = 0;
  auto mark = [&](const char *name)
    { if (!tl_on) return;
      cudaStreamSynchronize(c->st); cudaStreamSynchronize(c->cst);
      struct timespec t; clock_gettime(CLOCK_MONOTONIC,&t);
      const double ms = (t.tv_sec - tl_t0.tv_sec) * 1e3 + (t.tv_nsec - tl_t0.tv_nsec) * 1e-6;
      tl_len += snprintf(tl_buf + tl_len,sizeof(tl_buf) - (size_t) tl_len," %s %.2f",name,ms);
      tl_t0 = t;
    };

  /* ---- global position space: rank r's stream starts at pos_base[r] (multiples of 64) */
  std::vector<u64> all;
  { u64 mine = (u64) npos;
    int rc = mg_gather_u64(c,&mine,1,all);
    if (rc) return rc;
  }
  std::vector<u64> pbase(W + 1,0);
  for (int r = 0; r < W; r++) pbase[r+1] = pbase[r] + ((all[r] + 63) / 64) * 64;
  const long long total = (long long) pbase[W];
  /* the base strings are gathered by the SENDER, so a record only has to address its own rank's stream: the position field
     stays at the width of the longest local stream and the bits saved go to the bucket id (more, smaller buckets as the
     job grows: dedup and the on-chip tables see the same ~40 super-mers per bucket at 8 GPUs as at 1)                  */
  long long longest = 0;
  for (int r = 0; r < W; r++) longest = std::max<long long>(longest,(long long) (pbase[r+1] - pbase[r]));
  const SuperGeom g = super_geom(c->cfg.kmer,total,longest);
  if (g.pbits > 64 - SUP_LBITS - 2) return set_err(FKGPU_E_ARG,"multi-GPU count: %lld positions do not fit a super-mer record",longest);

  /* ---- scan own reads, level-1 partition by bucket */
  int rc = prepare_common(c,npos,std::max(g.P1,1),false,2);
  if (rc) return rc;
  if (c->segs.ensure(sizeof(SuperCounters))) return set_err(FKGPU_E_NOMEM,"out of device memory");
  SuperCounters *d_cnt = (SuperCounters *) c->segs.p;
  CU(cudaMemsetAsync(d_cnt,0,sizeof(SuperCounters),c->st));
  const size_t abytes = (size_t) (npos + 4) * 16;
  const u64 scap = (u64) (abytes / 2 / sizeof(u64)) - 8;
  Key<1> *SA = (Key<1> *) c->bufA.p;
  Key<1> *SB = (Key<1> *) ((char *) c->bufA.p + ((abytes / 2) & ~(size_t) 15));
  stage_begin(c,FKGPU_ST_SUPERSCAN);
  rc = super_scan_stage(c,d_seq,d_val,npos,g,0,(u64 *) SA,scap,d_cnt);
  if (rc) return rc;
  stage_end(c,FKGPU_ST_SUPERSCAN);
  SuperCounters hc;
  CU(cudaMemcpyAsync(&hc,d_cnt,sizeof(hc),cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));
  /* a rank whose super-mers overflow its staging buffer must not leave the others waiting in a collective: agree first */
  { u64 ok = (hc.nrec <= scap) ? 1 : 0;
    rc = mg_gather_u64(c,&ok,1,all);
    if (rc) return rc;
    for (int r = 0; r < W; r++)
      if (!all[r]) return set_err(FKGPU_E_UNSUPPORTED,"multi-GPU count: rank %d produced more super-mers than its staging buffer holds",r);
  }
  mark("scan");
  const long long S = (long long) hc.nrec;
  const int b1 = g.bbits ? g.P1 : 0, n1 = 1 << b1;
  stage_begin(c,FKGPU_ST_SUPERPART);
  rc = super_level1(c,SA,SB,S,b1);
  if (rc) return rc;
  stage_end(c,FKGPU_ST_SUPERPART);

  /* ---- owners of the buckets: all-reduced level-1 histogram -> contiguous ranges */
  if (m->small.ensure((size_t) (FKGPU_HIST_BINS + 2*n1 + 64) * 8)) return set_err(FKGPU_E_NOMEM,"out of device memory");
  std::vector<u64> gh(n1), off(n1 + 1);
  { u64 *d_g = (u64 *) m->small.p;
    NC(na->AllReduce(c->hist1.p,d_g,(size_t) n1,fkmg::kUint64,fkmg::kSum,m->comm,c->st));
    CU(cudaMemcpyAsync(gh.data(),d_g,(size_t) n1 * 8,cudaMemcpyDeviceToHost,c->st));
    CU(cudaMemcpyAsync(off.data(),c->off1.p,(size_t) (n1 + 1) * 8,cudaMemcpyDeviceToHost,c->st));
    CU(cudaStreamSynchronize(c->st));
  }
  std::vector<int> beg;
  fkmg::splitters(gh.data(),n1,W,beg);
  MgPlan pl;
  rc = mg_plan(c,off,beg,pl);
  if (rc) return rc;
  m->sent_records = (int64_t) (pl.nsend - pl.scnt[me]);
  mark("level1+plan");

  /* ---- the base string of every record travels beside it, compact: ceil((l + k - 1) / 16) words, back to back.
          Words per record -> exclusive scan -> strings + the records to send (position = word offset inside the slice)     */
  if (c->scnt.ensure((size_t) (S + 2) * 4) || c->poff.ensure((size_t) (S + 2) * 8))
    return set_err(FKGPU_E_NOMEM,"out of device memory (payload offsets)");
  Misc *d_misc0 = (Misc *) c->misc.p;
  u64 wtotal = 0;
  std::vector<u64> wcut(W + 1,0);
  stage_begin(c,FKGPU_ST_SCATTER);
  if (S > 0)
    { k_payload_words<<<(unsigned) ((S + 255) / 256),256,0,c->st>>>((const u64 *) SB,S,g.pbits,g.k,(u32 *) c->scnt.p); KCHECK(); }
  rc = run_large_scan<2>(c,(const u32 *) c->scnt.p,S,(u64 *) c->poff.p,&d_misc0->total_pass);
  if (rc) return rc;
  CU(cudaMemcpyAsync(&wtotal,&d_misc0->total_pass,8,cudaMemcpyDeviceToHost,c->st));
  for (int r = 0; r < W; r++)
    if (pl.soff[r] < (u64) S) CU(cudaMemcpyAsync(&wcut[r],(const u64 *) c->poff.p + pl.soff[r],8,cudaMemcpyDeviceToHost,c->st));
  CU(cudaStreamSynchronize(c->st));
  for (int r = 0; r < W; r++) if (pl.soff[r] >= (u64) S) wcut[r] = wtotal;
  wcut[W] = wtotal;
  /* what crosses, in words: the slices are contiguous and in rank order, so slice r = words [wcut[r], wcut[r+1]);
     an all-gather of the counts plans the receive side                                                                  */
  MgPlan pw;
  pw.soff.resize(W); pw.scnt.resize(W); pw.roff.resize(W); pw.rcnt.resize(W);
  for (int r = 0; r < W; r++) { pw.soff[r] = wcut[r]; pw.scnt[r] = wcut[r+1] - wcut[r]; }
  rc = mg_gather_u64(c,pw.scnt.data(),W,all);
  if (rc) return rc;
  pw.nrecv = 0;
  for (int r = 0; r < W; r++) { pw.rcnt[r] = all[(size_t) r * W + me]; pw.roff[r] = pw.nrecv; pw.nrecv += pw.rcnt[r]; }
  if (m->payload.ensure((size_t) (wtotal + 16) * 4) || m->rrec.ensure((size_t) (pl.nrecv + 8) * 8) || m->rscr.ensure((size_t) (pl.nrecv + 8) * 8)
      || m->rpay.ensure((size_t) (pw.nrecv + 16) * 4))
    return set_err(FKGPU_E_NOMEM,"out of device memory (exchange buffers: %lld records out, %llu in)",S,pl.nrecv);
  /* SA is free again (the scan's output went through the partition into SB): it takes the records to send */
  SliceTable sst; memset(&sst,0,sizeof(sst));
  sst.n = W;
  for (int r = 0; r < W; r++) sst.start[r] = pl.soff[r];      /* empty slices share their start with the next one: the last start <= i wins */
  if (S > 0)
    { k_materialise_compact<<<(unsigned) ((S + 255) / 256),256,0,c->st>>>((const u64 *) SB,S,g.pbits,0,g.k,d_seq,(const u64 *) c->poff.p,
                                                                        sst,(u64 *) SA,(u32 *) m->payload.p); KCHECK();
    }
  mark("payload-build");
  /* the records go first, on the compute stream; their base strings follow on the copy stream, so that the larger payload
     crosses NVLink while this rank already partitions the records it received (the counting kernel waits for ev_pay)      */
  rc = mg_alltoall(c,SA,pl.soff,pl.scnt,m->rrec.p,pl.roff,pl.rcnt,8,c->st);
  if (rc) return rc;
  stage_end(c,FKGPU_ST_SCATTER);
  CU(cudaEventRecord(m->ev_rec,c->st));
  CU(cudaStreamWaitEvent(c->cst,m->ev_rec,0));
  rc = mg_alltoall(c,m->payload.p,pw.soff,pw.scnt,m->rpay.p,pw.roff,pw.rcnt,4,c->cst);
  if (rc) return rc;
  CU(cudaEventRecord(m->ev_pay,c->cst));
  const long long nrecv = (long long) pl.nrecv;
  if (nrecv > 0)
    { SliceTable rst; memset(&rst,0,sizeof(rst));
      rst.n = W;
      for (int r = 0; r < W; r++) { rst.start[r] = pl.roff[r]; rst.add[r] = pw.roff[r]; }
      k_rebase_slices<<<(unsigned) ((nrecv + 255) / 256),256,0,c->st>>>((u64 *) m->rrec.p,nrecv,rst); KCHECK();
    }

  mark("exchange");
  /* ---- count the owned buckets; the entries are bounded by the k-mers the received records cover */
  u64 nk = 0;
  { u64 *d_nk = (u64 *) m->small.p;
    CU(cudaMemsetAsync(d_nk,0,8,c->st));
    if (nrecv > 0)
      { k_sum_lengths<<<c->sms * 4,256,0,c->st>>>((const u64 *) m->rrec.p,nrecv,g.pbits,d_nk); KCHECK(); }
    CU(cudaMemcpyAsync(&nk,d_nk,8,cudaMemcpyDeviceToHost,c->st));
    CU(cudaStreamSynchronize(c->st));
  }
  void *ent = NULL;
  if (want_entries)
    { if (c->bufB.ensure((size_t) (nk + 4) * EB)) return set_err(FKGPU_E_NOMEM,"out of device memory (entries of %llu k-mers)",nk);
      ent = c->bufB.p;
    }
  rc = prepare_small(c,std::max(g.P1,1));                 /* histogram + scalars start from zero; hist1 / off1 are re-used below */
  if (rc) return rc;
  if (c->segs.ensure(sizeof(SuperCounters))) return set_err(FKGPU_E_NOMEM,"out of device memory");
  d_cnt = (SuperCounters *) c->segs.p;
  CU(cudaMemsetAsync(d_cnt,0,sizeof(SuperCounters),c->st));
  Misc hm; long long gmax = 0;
  rc = super_count_stage(c,g,(Key<1> *) m->rrec.p,(Key<1> *) m->rscr.p,nrecv,NULL,1,NULL,NULL,m->rpay.p,(void *) m->ev_pay,ent,nk,d_cnt,&hc,&hm,&gmax);
  if (rc) return rc;
  const u64 nent = want_entries ? hc.nent : 0;
  bank_times(c);
  mark("count");

  /* ---- table: distinct entries to the owners of their key range, key order there */
  res->ntable = 0; res->table = NULL; res->table_dev = NULL;
  m->sent_entries = 0;
  if (want_entries)
    { const int eb1 = 11, ne = 1 << eb1;
      if (m->epart.ensure((size_t) (nent + 8) * EB) || m->small.ensure((size_t) (FKGPU_HIST_BINS + 4*ne + 64) * 8))
        return set_err(FKGPU_E_NOMEM,"out of device memory (entry partition)");
      u64 *d_h = (u64 *) m->small.p, *d_o = d_h + ne, *d_g = d_o + ne + 1;
      stage_begin(c,FKGPU_ST_L2HIST);
      rc = (EW == 3) ? entries_partition_t<3>(c,ent,(int64_t) nent,eb1,m->epart.p,(uint64_t *) d_h,(uint64_t *) d_o)
                     : entries_partition_t<2>(c,ent,(int64_t) nent,eb1,m->epart.p,(uint64_t *) d_h,(uint64_t *) d_o);
      if (rc) return rc;
      stage_end(c,FKGPU_ST_L2HIST);
      std::vector<u64> gh2(ne), off2(ne + 1);
      NC(na->AllReduce(d_h,d_g,(size_t) ne,fkmg::kUint64,fkmg::kSum,m->comm,c->st));
      CU(cudaMemcpyAsync(gh2.data(),d_g,(size_t) ne * 8,cudaMemcpyDeviceToHost,c->st));
      CU(cudaMemcpyAsync(off2.data(),d_o,(size_t) (ne + 1) * 8,cudaMemcpyDeviceToHost,c->st));
      CU(cudaStreamSynchronize(c->st));
      std::vector<int> beg2;
      fkmg::splitters(gh2.data(),ne,W,beg2);
      MgPlan p2;
      rc = mg_plan(c,off2,beg2,p2);
      if (rc) return rc;
      m->sent_entries = (int64_t) (p2.nsend - p2.scnt[me]);
      if (m->erecv.ensure((size_t) (p2.nrecv + 8) * EB) || c->bufA.ensure((size_t) (p2.nrecv + 8) * EB))
        return set_err(FKGPU_E_NOMEM,"out of device memory (%llu entries in)",p2.nrecv);
      mark("entry-partition+plan");
      rc = mg_alltoall(c,m->epart.p,p2.soff,p2.scnt,m->erecv.p,p2.roff,p2.rcnt,EB,c->st);
      if (rc) return rc;
      bank_times(c);
      mark("entry-exchange");
      fkgpu_result tr; memset(&tr,0,sizeof(tr));
      rc = entries_sort_stage(c,m->erecv.p,c->bufA.p,(long long) p2.nrecv,fetch_table,&tr);
      if (rc) return rc;
      rc = d2h_join(c);
      if (rc) return rc;
      res->ntable = tr.ntable; res->table = tr.table; res->table_dev = tr.table_dev;
      mark("entry-sort+d2h");
    }

  /* ---- global histogram and scalars; table sizes in rank (= key) order */
  { u64 *d_x = (u64 *) m->small.p, *d_s = d_x + FKGPU_HIST_BINS;
    u64 sc[3] = { hm.maxinst, nk, hm.ndistinct };
    CU(cudaMemcpyAsync(d_s,sc,sizeof(sc),cudaMemcpyHostToDevice,c->st));
    NC(na->AllReduce(c->ghist.p,d_x,(size_t) FKGPU_HIST_BINS,fkmg::kUint64,fkmg::kSum,m->comm,c->st));
    NC(na->AllReduce(d_s,d_s,3,fkmg::kUint64,fkmg::kSum,m->comm,c->st));
    CU(cudaMemcpyAsync(c->h_hist,d_x,sizeof(c->h_hist),cudaMemcpyDeviceToHost,c->st));
    CU(cudaMemcpyAsync(sc,d_s,sizeof(sc),cudaMemcpyDeviceToHost,c->st));
    CU(cudaStreamSynchronize(c->st));
    res->hist = c->h_hist;
    res->max_inst = (int64_t) sc[0]; res->nkmers = (int64_t) sc[1]; res->ndistinct = (int64_t) sc[2];
    u64 nt = (u64) res->ntable;
    rc = mg_gather_u64(c,&nt,1,all);
    if (rc) return rc;
    m->table_sizes.assign(W,0); m->global_ntable = 0;
    for (int r = 0; r < W; r++) { m->table_sizes[r] = (int64_t) all[r]; m->global_ntable += (int64_t) all[r]; }
  }
  mark("reduce");
  if (tl_on) fprintf(stderr,"[fkgpu mg rank %d]%s\n",me,tl_buf);
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES+1],c->st);
  CU(cudaStreamSynchronize(c->st));
  collect_times(c,res);
  single_run(c,res);
  c->last_path = 1; c->last_ndist = (long long) hm.ndistinct;
  c->st_super = S; c->st_ent = (long long) nent; c->st_groups = gmax; c->st_split = hc.pad; c->st_spill = c->st_spill_acc; c->st_expanded = c->st_exp_acc;
  return FKGPU_OK;
}

extern "C" int fkgpu_count_packed_multi(fkgpu_ctx *c, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos,
                                        int fetch_table, fkgpu_result *res)
{ if (c == NULL || res == NULL || npos < 0 || (npos > 0 && (d_seq == NULL || d_val == NULL)))
    return set_err(FKGPU_E_ARG,"fkgpu_count_packed_multi: bad argument");
  if (c->mg == NULL) return set_err(FKGPU_E_STATE,"fkgpu_count_packed_multi: no communicator (fkgpu_comm_init)");
  CU(cudaSetDevice(c->cfg.device));
  init_result(c,res);
  res->nbases = npos;
  return count_packed_multi(c,d_seq,d_val,npos,fetch_table,res);
}

/*  pieces: the counts of positions src[i] .. src[i]+len[i]) go to the output at dst[i]; offs = profile start of every read
 *  (+ total).  segs (optional): the stretches of the read stream in the order of the output -- stretch s covers positions
 *  [off, off+fill) and holds pieces [p0, p1), sorted by src -- which lets the lookup kernel write the output in place and in
 *  order (k_profile<.,.,true>).  Without segs the pieces may lie anywhere: one u16 per position, then a gather.           */
struct ProfSeg { long long off, fill; long long p0, p1; };
struct ProfPieces { std::vector<long long> src, dst; std::vector<int> len; std::vector<int64_t> offs; long long run = 0; std::vector<ProfSeg> segs; };

#define FKGPU_PROF_SLICES 16      /* an ordered profile run is launched and copied back in this many slices */

template<int NW, bool HASH>
static int profiles_launch(fkgpu_ctx *c, const ProfileParams &q, bool direct, long long ntiles)
{ const size_t sm = (size_t) (SCAN_SEQW + SCAN_VALW) * 4;
  if (ntiles <= 0) return FKGPU_OK;
  if (direct) k_profile<NW,HASH,true><<<(unsigned) ntiles,SCAN_TPB,sm,c->st>>>(q);
  else        k_profile<NW,HASH,false><<<(unsigned) ntiles,SCAN_TPB,sm,c->st>>>(q);
  KCHECK();
  return FKGPU_OK;
}

template<int NW>
static int profiles_run(fkgpu_ctx *c, const u32 *d_seq, const u32 *d_val, long long npos, const ProfPieces &pp,
                        int64_t *nreads, const int64_t **off, const uint16_t **prof)
{ const size_t np = pp.src.size();
  const long long run = pp.run;
  const bool hash = c->ph_on;
  const bool direct = !prof_legacy() && !pp.segs.empty() && np < (size_t) 0x7fffffff;
  if (c->pout.ensure((size_t) (run + 2) * 2) || c->psrc.ensure(np*8 + 8) || c->pdst.ensure(np*8 + 8) || c->plen.ensure(np*4 + 8)
      || (!direct && c->praw.ensure((size_t) (npos + 64) * 2)))
    return set_err(FKGPU_E_NOMEM,"out of device memory (profiles of %lld positions)",npos);
  if (c->h_prof.ensure((size_t) (run + 2) * 2) || c->h_poff.ensure(pp.offs.size() * 8))
    return set_err(FKGPU_E_NOMEM,"out of pinned host memory (profiles)");
  if (c->ev_prof == nullptr) CU(cudaEventCreateWithFlags(&c->ev_prof,cudaEventDisableTiming));
  stage_begin(c,FKGPU_ST_PROFILE);
  ProfileParams q;
  memset(&q,0,sizeof(q));
  ScanGeom g = scan_geom(c,npos,0);
  fill_scan_params(c,q.sp,d_seq,d_val,npos,0,g);
  q.keys = c->pkeys.p; q.cnts = (const uint16_t *) c->pcnts.p; q.idx = (const u64 *) c->pidx.p; q.B = c->ptab_B;
  q.H.slots = (const ulonglong2 *) c->phash.p; q.H.hcnt = (const uint16_t *) c->phcnt.p; q.H.nbuckets = c->ph_nbuckets; q.H.wide = c->ph_wide;
  { static int ltc = -1;            /* L2::64B on the lookup loads: DRAM bytes per lookup 127 -> 67 (ncu r2q), same time; FKGPU_PROF_LTC=0 drops the hint */
    if (ltc < 0) { const char *e = getenv("FKGPU_PROF_LTC"); ltc = e ? atoi(e) : 64; }
    if (ltc == 64) q.H.wide |= 2;
  }
  q.raw = (uint16_t *) c->praw.p;
  q.psrc = (const long long *) c->psrc.p; q.pdst = (const long long *) c->pdst.p; q.plen = (const int *) c->plen.p;
  q.out = (uint16_t *) c->pout.p;
  if (np > 0)
    { CU(cudaMemcpyAsync(c->psrc.p,pp.src.data(),np*8,cudaMemcpyHostToDevice,c->st));
      CU(cudaMemcpyAsync(c->pdst.p,pp.dst.data(),np*8,cudaMemcpyHostToDevice,c->st));
      CU(cudaMemcpyAsync(c->plen.p,pp.len.data(),np*4,cudaMemcpyHostToDevice,c->st));
    }
  int rc = FKGPU_OK;
  if (!direct)
    { rc = hash ? profiles_launch<NW,true>(c,q,false,g.ntiles) : profiles_launch<NW,false>(c,q,false,g.ntiles);
      if (rc) return rc;
      if (np > 0)
        { k_gather_profile<<<c->sms * 8,256,0,c->st>>>((const uint16_t *) c->praw.p,q.psrc,q.pdst,q.plen,(long long) np,(uint16_t *) c->pout.p); KCHECK(); }
      stage_end(c,FKGPU_ST_PROFILE);
      if (run > 0) CU(cudaMemcpyAsync(c->h_prof.p,c->pout.p,(size_t) run * 2,cudaMemcpyDeviceToHost,c->st));
    }
  else
    { /* virtual tiles: the stretches cut into SCAN_TILE positions, with the pieces that meet each tile and the first output
         index the tile (or a later one) writes                                                                          */
      std::vector<long long> vsrc, vdst;
      std::vector<int> vlo, vend;
      for (const ProfSeg &sg : pp.segs)
        { long long a = sg.p0, b = sg.p0;
          for (long long P = sg.off; P < sg.off + sg.fill; P += SCAN_TILE)
            { while (a < sg.p1 && pp.src[a] + pp.len[a] <= P) a++;
              while (b < sg.p1 && pp.src[b] < P + SCAN_TILE) b++;
              vsrc.push_back(P); vlo.push_back((int) a); vend.push_back((int) b);
              long long d;
              if (a < sg.p1) d = (pp.src[a] < P) ? pp.dst[a] + (P - pp.src[a]) : pp.dst[a];
              else           d = ((size_t) sg.p1 < np) ? pp.dst[sg.p1] : run;
              vdst.push_back(d);
            }
        }
      const long long nt = (long long) vsrc.size();
      if (c->vtsrc.ensure((size_t) nt * 8 + 8) || c->vtplo.ensure((size_t) nt * 4 + 8) || c->vtpend.ensure((size_t) nt * 4 + 8))
        return set_err(FKGPU_E_NOMEM,"out of device memory (profile tiles)");
      if (nt > 0)
        { CU(cudaMemcpyAsync(c->vtsrc.p,vsrc.data(),(size_t) nt * 8,cudaMemcpyHostToDevice,c->st));
          CU(cudaMemcpyAsync(c->vtplo.p,vlo.data(),(size_t) nt * 4,cudaMemcpyHostToDevice,c->st));
          CU(cudaMemcpyAsync(c->vtpend.p,vend.data(),(size_t) nt * 4,cudaMemcpyHostToDevice,c->st));
        }
      q.vt_src = (const long long *) c->vtsrc.p; q.vt_plo = (const int *) c->vtplo.p; q.vt_pend = (const int *) c->vtpend.p;
      for (int sl = 0; sl < FKGPU_PROF_SLICES; sl++)
        { const long long t0 = nt * sl / FKGPU_PROF_SLICES, t1 = nt * (sl + 1) / FKGPU_PROF_SLICES;
          if (t1 <= t0) continue;
          q.vt0 = t0;
          rc = hash ? profiles_launch<NW,true>(c,q,true,t1 - t0) : profiles_launch<NW,false>(c,q,true,t1 - t0);
          if (rc) return rc;
          const long long d0 = vdst[t0], d1 = (t1 < nt) ? vdst[t1] : run;
          if (d1 > d0)
            { CU(cudaEventRecord(c->ev_prof,c->st));
              CU(cudaStreamWaitEvent(c->cst,c->ev_prof,0));
              CU(cudaMemcpyAsync((uint16_t *) c->h_prof.p + d0,(const uint16_t *) c->pout.p + d0,(size_t) (d1 - d0) * 2,cudaMemcpyDeviceToHost,c->cst));
            }
        }
      stage_end(c,FKGPU_ST_PROFILE);
      CU(cudaStreamSynchronize(c->cst));
    }
  CU(cudaStreamSynchronize(c->st));
  cudaEventElapsedTime(&c->ms[FKGPU_ST_PROFILE],c->ev[2*FKGPU_ST_PROFILE],c->ev[2*FKGPU_ST_PROFILE+1]);
  memcpy(c->h_poff.p,pp.offs.data(),pp.offs.size()*8);
  *nreads = (int64_t) pp.offs.size() - 1;
  *off = (const int64_t *) c->h_poff.p;
  *prof = (const uint16_t *) c->h_prof.p;
  return FKGPU_OK;
}

template<int NW>
static int profiles_t(fkgpu_ctx *c, int64_t *nreads, const int64_t **off, const uint16_t **prof)
{ const int k = c->cfg.kmer, bc = c->cfg.bc_prefix;
  /* pieces in tid-major order; a continuation piece (rem carry) extends the previous read */
  ProfPieces pp;
  bool ordered = true;
  for (auto &t : c->tids)
    { size_t ch = 0;
      long long last = -1;
      for (size_t i = 0; i < t.rstart.size(); i++)
        { const int b = t.rcont[i] ? 0 : bc;
          int pl = t.rlen[i] - b - k + 1; if (pl < 0) pl = 0;
          if (!t.rcont[i]) pp.offs.push_back(pp.run);
          /* the chunk of this thread that holds the piece: chunks and pieces both come in arrival order */
          while (ordered && ch < t.chunks.size() && !(t.rstart[i] >= t.chunks[ch].first && t.rstart[i] < t.chunks[ch].first + t.chunks[ch].second))
            { ch++; last = -1; }
          if (ordered && (ch >= t.chunks.size() || t.rstart[i] < last)) ordered = false;
          if (ordered)
            { if (pp.segs.empty() || pp.segs.back().off != t.chunks[ch].first || last < 0)
                { ProfSeg sg; sg.off = t.chunks[ch].first; sg.fill = t.chunks[ch].second; sg.p0 = sg.p1 = (long long) pp.src.size();
                  pp.segs.push_back(sg);
                }
              pp.segs.back().p1 = (long long) pp.src.size() + 1;
              last = t.rstart[i] + t.rlen[i];
            }
          pp.src.push_back(t.rstart[i] + b); pp.dst.push_back(pp.run); pp.len.push_back(pl);
          pp.run += pl;
        }
    }
  pp.offs.push_back(pp.run);
  if (!ordered) pp.segs.clear();
  return profiles_run<NW>(c,(const u32 *) c->seq.p,(const u32 *) c->val.p,c->last_npos,pp,nreads,off,prof);
}

template<int NW>
static int load_profile_table_t(fkgpu_ctx *c, const uint8_t *records, int64_t n)
{ const int tw = c->kbytes + 2;
  const u64 U = (u64) n;
  if (c->table.ensure((size_t) U * tw + 64) || c->pkeys.ensure((size_t) (U + 1) * sizeof(Key<NW>)) || c->pcnts.ensure((size_t) (U + 1) * 2))
    return set_err(FKGPU_E_NOMEM,"out of device memory (profile lookup table of %lld k-mers)",(long long) n);
  if (U > 0)
    { CU(cudaMemcpyAsync(c->table.p,records,(size_t) U * tw,cudaMemcpyHostToDevice,c->st));
      k_records_to_keys<NW><<<(unsigned) ((U + 255) / 256),256,0,c->st>>>((const uint8_t *) c->table.p,U,c->kbytes,(Key<NW> *) c->pkeys.p,(uint16_t *) c->pcnts.p); KCHECK();
    }
  { int rc = build_profile_lookup<NW>(c,U);
    if (rc) return rc;
  }
  CU(cudaStreamSynchronize(c->st));
  c->ptab_n = (long long) U; c->res_nw = NW; c->rel_table = true;
  return FKGPU_OK;
}

extern "C" int fkgpu_load_profile_table(fkgpu_ctx *c, const uint8_t *records, int64_t n)
{ if (c == NULL || n < 0 || (n > 0 && records == NULL)) return set_err(FKGPU_E_ARG,"fkgpu_load_profile_table: bad argument");
  if (!c->cfg.do_profile) return set_err(FKGPU_E_STATE,"fkgpu_load_profile_table: the context was created without do_profile");
  CU(cudaSetDevice(c->cfg.device));
  const int tw = c->kbytes + 2;
  for (int64_t i = 1; i < n; i++)                      /* the lookup is a binary search: the records must be strictly increasing */
    if (memcmp(records + (i-1)*tw,records + i*tw,(size_t) c->kbytes) >= 0)
      return set_err(FKGPU_E_ARG,"fkgpu_load_profile_table: records %lld and %lld are not in increasing key order",(long long) (i-1),(long long) i);
  return (c->NW == 1) ? load_profile_table_t<1>(c,records,n) : load_profile_table_t<2>(c,records,n);
}

/*  GPU Fastmerge (Fastmerge.c:168-450): the records of ntab sorted tables are put in key order together by the
 *  entry sort (two-level prefix partition + in-smem ordering, the path every count ends with), then runs of equal k-mers are
 *  merged: counts added and saturated, histogram of the merged counts.  res->max_inst only holds what the unsaturated members
 *  of saturated sums stood for (Fastmerge.c:321-327): the caller adds the max_inst of the input histograms (Fastmerge.c:1009). */
extern "C" int fkgpu_merge_tables(fkgpu_ctx *c, const uint8_t *const *tables, const int64_t *n, int ntab, int fetch_table, fkgpu_result *res)
{ if (c == NULL || res == NULL || ntab < 1 || tables == NULL || n == NULL) return set_err(FKGPU_E_ARG,"fkgpu_merge_tables: bad argument");
  if (c->cfg.do_table < 1 || c->cfg.do_profile) return set_err(FKGPU_E_STATE,"fkgpu_merge_tables: needs a context with do_table >= 1 and no do_profile");
  CU(cudaSetDevice(c->cfg.device));
  init_result(c,res);
  const int tw = c->kbytes + 2, EW = entry_words(c->cfg.kmer);
  const size_t EB = (size_t) 8 * EW;
  long long N = 0;
  for (int t = 0; t < ntab; t++)
    { if (n[t] < 0 || (n[t] > 0 && tables[t] == NULL)) return set_err(FKGPU_E_ARG,"fkgpu_merge_tables: table %d is missing",t);
      N += n[t];
    }
  int rc = prepare_small(c,1);
  if (rc) return rc;
  if (c->spillA.ensure((size_t) N * tw + 64) || c->bufB.ensure((size_t) (N + 4) * EB) || c->bufA.ensure((size_t) (N + 4) * EB)
      || c->scnt.ensure((size_t) (N + 2) * 4) || c->poff.ensure((size_t) (N + 2) * 8))
    return set_err(FKGPU_E_NOMEM,"out of device memory (merging %lld table records)",N);
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES],c->st);
  { size_t at = 0;
    for (int t = 0; t < ntab; t++)
      if (n[t] > 0)
        { CU(cudaMemcpyAsync((uint8_t *) c->spillA.p + at,tables[t],(size_t) n[t] * tw,cudaMemcpyHostToDevice,c->st));
          at += (size_t) n[t] * tw;
        }
  }
  res->ntable = 0; res->table = NULL; res->table_dev = NULL;
  if (N > 0)
    { const unsigned gr = (unsigned) ((N + 255) / 256);
      if (EW == 3) k_table_to_entries<3><<<gr,256,0,c->st>>>((const uint8_t *) c->spillA.p,(u64) N,c->kbytes,(Key<3> *) c->bufB.p,0,(u64) N);
      else         k_table_to_entries<2><<<gr,256,0,c->st>>>((const uint8_t *) c->spillA.p,(u64) N,c->kbytes,(Key<2> *) c->bufB.p,0,(u64) N);
      KCHECK();
      fkgpu_result tmp; memset(&tmp,0,sizeof(tmp));
      rc = entries_sort_stage(c,c->bufB.p,c->bufA.p,N,0,&tmp);            /* -> c->table: N records in key order, equal keys adjacent */
      if (rc) return rc;
      if (tmp.ntable != N) return set_err(FKGPU_E_CUDA,"internal: the merge sort returned %lld of %lld records",(long long) tmp.ntable,N);
      Misc *d_misc = (Misc *) c->misc.p;
      u32 *head = (u32 *) c->scnt.p;
      k_run_heads<<<gr,256,0,c->st>>>((const uint8_t *) c->table.p,(u64) N,c->kbytes,head); KCHECK();
      rc = run_large_scan<2>(c,head,N,(u64 *) c->poff.p,&d_misc->total_pass);
      if (rc) return rc;
      u64 U = 0;
      CU(cudaMemcpyAsync(&U,&d_misc->total_pass,8,cudaMemcpyDeviceToHost,c->st));
      CU(cudaStreamSynchronize(c->st));
      if (c->spillB.ensure((size_t) U * tw + 64)) return set_err(FKGPU_E_NOMEM,"out of device memory (merged table of %llu records)",U);
      CU(cudaMemsetAsync(c->ghist.p,0,FKGPU_HIST_BINS * 8,c->st));
      CU(cudaMemsetAsync(&d_misc->maxinst,0,8,c->st));
      k_merge_runs<<<gr,256,0,c->st>>>((const uint8_t *) c->table.p,(u64) N,c->kbytes,head,(const u64 *) c->poff.p,(uint8_t *) c->spillB.p,
                                       (u64 *) c->ghist.p,&d_misc->maxinst); KCHECK();
      res->ntable = (int64_t) U; res->table_dev = (const uint8_t *) c->spillB.p;
      if (fetch_table)
        { if (c->h_table.ensure((size_t) U * tw + 64)) return set_err(FKGPU_E_NOMEM,"out of pinned host memory (merged table)");
          CU(cudaMemcpyAsync(c->h_table.p,c->spillB.p,(size_t) U * tw,cudaMemcpyDeviceToHost,c->st));
          res->table = (const uint8_t *) c->h_table.p;
        }
    }
  Misc hm;
  CU(cudaMemcpyAsync(c->h_hist,c->ghist.p,sizeof(c->h_hist),cudaMemcpyDeviceToHost,c->st));
  CU(cudaMemcpyAsync(&hm,c->misc.p,sizeof(hm),cudaMemcpyDeviceToHost,c->st));
  cudaEventRecord(c->ev[2*FKGPU_NSTAGES+1],c->st);
  CU(cudaStreamSynchronize(c->st));
  collect_times(c,res);
  res->hist = c->h_hist; res->max_inst = (int64_t) hm.maxinst; res->ndistinct = res->ntable; res->nkmers = 0;
  single_run(c,res);
  return FKGPU_OK;
}

extern "C" int fkgpu_profiles_packed(fkgpu_ctx *c, const uint32_t *d_seq, const uint32_t *d_val, int64_t npos,
                                     const int64_t *read_start, const int32_t *read_len, int64_t nreads_in,
                                     int64_t *nreads, const int64_t **off, const uint16_t **prof)
{ if (c == NULL || nreads == NULL || off == NULL || prof == NULL || nreads_in < 0 || (nreads_in > 0 && (read_start == NULL || read_len == NULL))
      || (npos > 0 && (d_seq == NULL || d_val == NULL)))
    return set_err(FKGPU_E_ARG,"fkgpu_profiles_packed: bad argument");
  if (!c->cfg.do_profile) return set_err(FKGPU_E_STATE,"fkgpu_profiles_packed: the context was created without do_profile");
  if (c->ptab_n <= 0 && c->last_ndist > 0) return set_err(FKGPU_E_STATE,"fkgpu_profiles_packed: call fkgpu_count_packed first");
  CU(cudaSetDevice(c->cfg.device));
  const int k = c->cfg.kmer, bc = c->cfg.bc_prefix;
  ProfPieces pp;
  pp.src.reserve((size_t) nreads_in); pp.dst.reserve((size_t) nreads_in); pp.len.reserve((size_t) nreads_in); pp.offs.reserve((size_t) nreads_in + 1);
  for (int64_t i = 0; i < nreads_in; i++)
    { if (read_start[i] < 0 || read_len[i] < 0 || read_start[i] + read_len[i] > npos)
        return set_err(FKGPU_E_ARG,"fkgpu_profiles_packed: read %lld lies outside the stream",(long long) i);
      int pl = read_len[i] - bc - k + 1; if (pl < 0) pl = 0;
      pp.offs.push_back(pp.run);
      pp.src.push_back(read_start[i] + bc); pp.dst.push_back(pp.run); pp.len.push_back(pl);
      pp.run += pl;
    }
  pp.offs.push_back(pp.run);
  bool ordered = true;                 /* reads in stream order, not overlapping: the output can be written in place, in order */
  for (int64_t i = 1; i < nreads_in && ordered; i++) ordered = (read_start[i] >= read_start[i-1] + read_len[i-1]);
  if (ordered && nreads_in > 0)
    { ProfSeg sg; sg.off = 0; sg.fill = npos; sg.p0 = 0; sg.p1 = (long long) nreads_in;
      pp.segs.push_back(sg);
    }
  return (c->res_nw == 1) ? profiles_run<1>(c,d_seq,d_val,npos,pp,nreads,off,prof) : profiles_run<2>(c,d_seq,d_val,npos,pp,nreads,off,prof);
}

extern "C" int fkgpu_profiles(fkgpu_ctx *c, int64_t *nreads, const int64_t **off, const uint16_t **prof)
{ if (c == NULL || nreads == NULL || off == NULL || prof == NULL) return set_err(FKGPU_E_ARG,"fkgpu_profiles: NULL argument");
  if (!c->cfg.do_profile) return set_err(FKGPU_E_STATE,"fkgpu_profiles: the context was created without do_profile");
  if (!c->finished) return set_err(FKGPU_E_STATE,"fkgpu_profiles: call fkgpu_finish first");
  CU(cudaSetDevice(c->cfg.device));
  return (c->res_nw == 1) ? profiles_t<1>(c,nreads,off,prof) : profiles_t<2>(c,nreads,off,prof);
}
