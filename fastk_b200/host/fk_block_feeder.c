/*  fk_block_feeder.c -- the producer side of the boundary for measurements: NTHREADS pthreads, each handing ITS DATA_BLOCKs
 *  to fkgpu_ingest one at a time, exactly the calling pattern of the reference's input module (io.c:2659-2691 starts the
 *  threads, io.c:535,753 call Distribute_Block once per filled block and reuse the block at once).  The blocks are slices
 *  of one host array of fixed-length reads (bench.py's synthetic batch): `row_bytes` bytes per read including the
 *  terminator, `rows_per_block` reads per block.  Nothing here touches the device; the library does.
 *
 *  bench.py's e2e arm calls fk_feed_blocks through ctypes so that the timed region holds C threads, as a FastK host has,
 *  and not Python threads taking turns on the interpreter lock.                                                      */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include "fastk_gpu.h"

typedef struct
  { fkgpu_ctx     *ctx;
    const char    *base;
    const int32_t *boff;
    int64_t        nreads, nblocks;
    int32_t        row_bytes, rows_per_block;
    int            tid, nthr, contig, rc;
  } Feed;

static void *feeder(void *v)
{ Feed *f = (Feed *) v;
  int64_t b, beg, end, step;
  if (f->contig)            /* -p: thread order is file order, every thread owns a contiguous range (io.c's file partition) */
    { beg = f->nblocks * f->tid / f->nthr; end = f->nblocks * (f->tid + 1) / f->nthr; step = 1; }
  else
    { beg = f->tid; end = f->nblocks; step = f->nthr; }
  for (b = beg; b < end; b += step)
    { int64_t r0 = b * f->rows_per_block;
      int64_t nr = f->nreads - r0 < f->rows_per_block ? f->nreads - r0 : f->rows_per_block;
      f->rc = fkgpu_ingest(f->ctx,f->tid,f->base + r0 * f->row_bytes,f->boff,(int32_t) nr,0);
      if (f->rc != 0) break;
    }
  return NULL;
}

/*  -> 0, or the first failing fkgpu_ingest code  */
int fk_feed_blocks(fkgpu_ctx *ctx, int nthr, const char *base, int64_t nreads, int32_t row_bytes, int32_t rows_per_block, int contig)
{ if (ctx == NULL || base == NULL || nthr < 1 || nthr > 256 || row_bytes < 1 || rows_per_block < 1 || nreads < 0) return FKGPU_E_ARG;
  int32_t  *boff = (int32_t *) malloc(sizeof(int32_t) * ((size_t) rows_per_block + 1));
  Feed      f[256];
  pthread_t th[256];
  int       t, started = 0, rc = 0;
  if (boff == NULL) return FKGPU_E_NOMEM;
  for (t = 0; t <= rows_per_block; t++) boff[t] = t * row_bytes;
  for (t = 0; t < nthr; t++)
    { f[t].ctx = ctx; f[t].base = base; f[t].boff = boff; f[t].nreads = nreads;
      f[t].nblocks = (nreads + rows_per_block - 1) / rows_per_block;
      f[t].row_bytes = row_bytes; f[t].rows_per_block = rows_per_block;
      f[t].tid = t; f[t].nthr = nthr; f[t].contig = contig; f[t].rc = 0;
      if (pthread_create(th + t,NULL,feeder,f + t) != 0) { rc = FKGPU_E_NOMEM; break; }
      started += 1;
    }
  for (t = 0; t < started; t++)
    { pthread_join(th[t],NULL);
      if (rc == 0 && f[t].rc != 0) rc = f[t].rc;
    }
  free(boff);
  return rc;
}
