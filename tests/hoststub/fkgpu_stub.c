/* TEST INFRASTRUCTURE ONLY -- a recording stand-in for libfastk_gpu.so, used by tests/test_host_reader.py to check the
 * C host program's input side (file discovery, byte-range split over reader threads, FASTA/FASTQ automaton, gzip,
 * -c compression, DATA_BLOCK assembly with the rem / k-1 overlap convention of io.c:296-333) on a machine without a
 * GPU.  It counts nothing: fkgpu_ingest stores the reads it is handed, fkgpu_finish writes them, tid-major, one per
 * line, to $FKSTUB_OUT and returns an empty result.  Never linked into the product.                              */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "fastk_gpu.h"

typedef struct { char **reads; long long n, cap; int carry; long long blocks, maxblock, maxreads; } Tid;
struct fkgpu_ctx { fkgpu_config cfg; Tid *t; };
static int64_t HIST[FKGPU_HIST_BINS];
static char ERR[256];

const char *fkgpu_last_error(void) { return ERR; }
int fkgpu_device_count(void) { return 1; }

int fkgpu_create(const fkgpu_config *cfg, fkgpu_ctx **out)
{ fkgpu_ctx *c = (fkgpu_ctx *) calloc(1,sizeof(*c));
  c->cfg = *cfg;
  c->t = (Tid *) calloc(cfg->nthreads > 0 ? cfg->nthreads : 1,sizeof(Tid));
  *out = c;
  return 0;
}

void fkgpu_destroy(fkgpu_ctx *c) { (void) c; }

int fkgpu_ingest(fkgpu_ctx *c, int tid, const char *bases, const int32_t *boff, int32_t nreads, int32_t rem)
{ Tid *t = c->t + tid;
  int i;
  if (tid < 0 || tid >= c->cfg.nthreads) { snprintf(ERR,sizeof(ERR),"stub: bad tid %d",tid); return FKGPU_E_ARG; }
  t->blocks += 1;
  if (getenv("FKSTUB_DISCARD")) return 0;             /* reader throughput measurements: take the block and drop it */
  if (boff[nreads] - boff[0] > t->maxblock) t->maxblock = boff[nreads] - boff[0];
  if (nreads > t->maxreads) t->maxreads = nreads;
  for (i = 0; i < nreads; i++)
    { const char *s = bases + boff[i];
      long long len = boff[i+1] - boff[i] - 1;
      if (s[len] != '\0') { snprintf(ERR,sizeof(ERR),"stub: read %d of a block is not 0-terminated",i); return FKGPU_E_ARG; }
      if (i == 0 && t->carry)                       /* continuation: re-delivers the last k-1 bases of the previous piece */
        { char *prev = t->reads[t->n-1];
          long long pl = (long long) strlen(prev), ov = c->cfg.kmer - 1;
          if (len < ov || pl < ov || memcmp(prev + pl - ov,s,(size_t) ov) != 0)
            { snprintf(ERR,sizeof(ERR),"stub: continuation piece does not start with the k-1 overlap"); return FKGPU_E_ARG; }
          prev = (char *) realloc(prev,(size_t) (pl + len - ov + 1));
          memcpy(prev + pl,s + ov,(size_t) (len - ov));
          prev[pl + len - ov] = '\0';
          t->reads[t->n-1] = prev;
          continue;
        }
      if (t->n >= t->cap) { t->cap = 2*t->cap + 1024; t->reads = (char **) realloc(t->reads,sizeof(char *)*(size_t) t->cap); }
      t->reads[t->n] = (char *) malloc((size_t) len + 1);
      memcpy(t->reads[t->n],s,(size_t) len + 1);
      t->n += 1;
    }
  t->carry = (rem > 0);
  return 0;
}

int fkgpu_finish(fkgpu_ctx *c, int fetch_table, fkgpu_result *res)
{ const char *path = getenv("FKSTUB_OUT");
  FILE *f = path ? fopen(path,"w") : NULL;
  int tid; long long i, nr = 0, nb = 0;
  (void) fetch_table;
  for (tid = 0; tid < c->cfg.nthreads; tid++)
    { Tid *t = c->t + tid;
      if (f) fprintf(f,"#tid %d blocks %lld maxblock %lld maxreads %lld\n",tid,t->blocks,t->maxblock,t->maxreads);
      for (i = 0; i < t->n; i++)
        { if (f) { fputs(t->reads[i],f); fputc('\n',f); }
          nr += 1; nb += (long long) strlen(t->reads[i]);
        }
    }
  if (f) fclose(f);
  memset(res,0,sizeof(*res));
  res->kmer = c->cfg.kmer; res->kmer_bytes = (2*c->cfg.kmer + 7) >> 3;
  res->nreads = nr; res->nbases = nb; res->hist = HIST;
  return 0;
}

int fkgpu_profiles(fkgpu_ctx *c, int64_t *nreads, const int64_t **off, const uint16_t **prof)
{ (void) c; (void) nreads; (void) off; (void) prof; snprintf(ERR,sizeof(ERR),"stub: no profiles"); return FKGPU_E_STATE; }
int fkgpu_read_counts(fkgpu_ctx *c, int64_t *per_tid)
{ int t; for (t = 0; t < c->cfg.nthreads; t++) per_tid[t] = c->t[t].n; return 0; }

int fkgpu_load_profile_table(fkgpu_ctx *c, const uint8_t *records, int64_t n)
{ (void) c; (void) records; (void) n; return 0; }
int fkgpu_comm_id(uint8_t *id) { (void) id; return -6; }
int fkgpu_comm_init(fkgpu_ctx *c, int nranks, int rank, const uint8_t *id) { (void) c; (void) nranks; (void) rank; (void) id; return -6; }
