/*  fastk_shim.c -- the drop-in boundary, reference side.
 *
 *  Compiled TOGETHER WITH THE REFERENCE'S OWN, UNMODIFIED  FastK.c io.c table.c libfastk.c  (from where they
 *  lie; nothing is copied), this file supplies every symbol of the five source files it replaces --
 *  split.c, count.c, MSDsort.c, LSDsort.c, merge.c -- on top of the C ABI of libfastk_gpu.so:
 *
 *      reference symbol (FastK.h:117-131)            here
 *      int  Determine_Scheme(DATA_BLOCK *)           forces NPARTS = 1 (no minimizer scheme is needed)
 *      void Split_Kmers(Input_Partition *, char *)   fkgpu_create + Scan_All_Input (io.c:2659)
 *      void Distribute_Block(DATA_BLOCK *, int)      fkgpu_ingest                    (callback of io.c:535,753)
 *      void Sorting(char *path, char *root)          fkgpu_finish + <root>.hist + the L-files table.c:382-394 merges
 *      void Merge_Profiles(char *path, char *root)   fkgpu_profiles + .prof/.pidx writers
 *      void Split_Table(char *root)                  -p:<table> is not supported: error exit
 *      uint8 Comp[256], int64 *NUM_RID               globals the replaced files owned (count.c:58, split.c:1405)
 *
 *  Errors follow the reference: message on stderr, then Clean_Exit(1) (FastK.c:181-221).
 *  Built by fastk_b200/host/Makefile into integration/_build/FastK_refhost when /root/reference is present.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/resource.h>

#include "libfastk.h"      /* the reference's own headers (-I<reference dir>) */
#include "FastK.h"

#include "fastk_gpu.h"
#include "fk_files.h"

extern char *Prog_Name;

uint8  Comp[256];
int64 *NUM_RID;

static fkgpu_ctx   *CTX;
static fkgpu_result RES;
static int64        EST_POSITIONS;   /* estimate of the input size from the training block (as FastK.c:428 does) */

static void fail(const char *what)
{ fprintf(stderr,"\n%s: %s: %s\n",Prog_Name,what,fkgpu_last_error());
  Clean_Exit(1);
}

int Determine_Scheme(DATA_BLOCK *block)
{ NPARTS = 1;                       /* one in-HBM batch: table.c's merge degenerates to a copy */
  /* bases + one terminator per read, scaled from the portion read to the whole data set; only a hint: it lets the
     library size its buffers up front and pack + scan every chunk while io.c is still reading (an underestimate
     just ends that overlap early)                                                                               */
  EST_POSITIONS = (int64) ((block->totlen + (double) block->nreads) * block->ratio * 1.05) + (1 << 20);
  if (VERBOSE)
    fprintf(stderr,"  GPU path: canonical-prefix buckets on the device, no minimizer scheme\n");
  return (KMER > 5 ? KMER-4 : 1);   /* MAX_SUPER: only sizes fields this path never uses */
}

void Split_Kmers(Input_Partition *io, char *root)
{ fkgpu_config cfg;
  (void) root;
  if (VERBOSE)
    fprintf(stderr,"\nPhase 1: Streaming the input to the GPU k-mer counter\n");
  /* FastK.c:473-489 has just lowered RLIMIT_NOFILE to (NPARTS+3)*NTHREADS+tid, sized for the temp files of the
     replaced stages; the CUDA driver needs its own descriptors, so lift the soft limit back to the hard one. */
  { struct rlimit rlp;
    if (getrlimit(RLIMIT_NOFILE,&rlp) == 0)
      { rlp.rlim_cur = rlp.rlim_max;
        setrlimit(RLIMIT_NOFILE,&rlp);
      }
  }
  memset(&cfg,0,sizeof(cfg));
  cfg.kmer = KMER; cfg.do_table = DO_TABLE; cfg.do_profile = DO_PROFILE; cfg.bc_prefix = BC_PREFIX;
  cfg.device = getenv("FASTK_GPU") ? atoi(getenv("FASTK_GPU")) : 0;
  cfg.nthreads = ITHREADS;
  cfg.reserve_bases = EST_POSITIONS;
  cfg.mem_limit = getenv("FASTK_GPU_MEM_GB") ? (int64_t) (atof(getenv("FASTK_GPU_MEM_GB")) * 1073741824.) : 0;
  if (fkgpu_create(&cfg,&CTX) != 0) fail("fkgpu_create");
  if (PRO_TABLE != NULL)
    { /* -p:<table>: what Split_Table (split.c:1943-2131) does for the reference's merge join -- bring the table to where the
         profiles are made -- done here, before the reads stream in, through the reference's own table reader               */
      Kmer_Stream *S = PRO_TABLE;
      uint8 *rec = (uint8 *) malloc((size_t) (S->nels > 0 ? S->nels : 1) * S->tbyte);
      int64  i = 0;
      if (rec == NULL) { fprintf(stderr,"%s: Out of memory (loading the -p table)\n",Prog_Name); Clean_Exit(1); }
      for (First_Kmer_Entry(S); S->csuf != NULL && i < S->nels; Next_Kmer_Entry(S))
        Current_Entry(S,rec + (i++)*S->tbyte);
      if (fkgpu_load_profile_table(CTX,rec,i) != 0) fail("fkgpu_load_profile_table");
      free(rec);
    }
  NUM_RID = (int64 *) calloc(ITHREADS > 0 ? ITHREADS : 1,sizeof(int64));
  Scan_All_Input(io);
}

void Distribute_Block(DATA_BLOCK *block, int tid)
{ if (fkgpu_ingest(CTX,tid,block->bases,block->boff,block->nreads,block->rem) != 0)
    fail("fkgpu_ingest");
}

void Split_Table(char *root)
{ (void) root;            /* the table went to the device in Split_Kmers: nothing is split into minimizer parts here */
  if (VERBOSE)
    fprintf(stderr,"\n  Profiles will be relative to %s (%lld %d-mers on the device)\n",PRO_NAME,(long long) PRO_TABLE->nels,KMER);
}

void Sorting(char *path, char *root)
{ if (VERBOSE)
    fprintf(stderr,"\nPhase 2: Sorting & Counting K-mers on the GPU\n");
  if (fkgpu_finish(CTX,DO_TABLE > 0,&RES) != 0) fail("fkgpu_finish");
  if (VERBOSE)
    fprintf(stderr,"  %lld %d-mers, %lld distinct, %.3f ms on the device\n",
                   (long long) RES.nkmers,KMER,(long long) RES.ndistinct,RES.ms_total);

  if (PRO_TABLE == NULL && fk_write_hist(path,root,KMER,RES.hist,RES.max_inst))
    { fprintf(stderr,"%s: Cannot write to %s/%s.hist.  Enough disk space?\n",Prog_Name,path,root);
      Clean_Exit(1);
    }

  if (DO_TABLE > 0)            /* what table_write_thread leaves for Merge_Tables (count.c:564-616,1560-1626) */
    { int   *beg = (int *) malloc(sizeof(int)*(NTHREADS+1));
      char  *name = (char *) malloc(strlen(SORT_PATH) + strlen(root) + 100);
      int    t, n;
      /* one sorted run per round of the count = one "part" of the reference: Merge_Tables merges <root>.<n>.L<t> over
         n < NPARTS for every thread t (table.c:382-394); the first-byte ranges of the threads are the same in every part */
      NPARTS = RES.nruns;
      IDX_BYTES = fk_idx_bytes(RES.ntable,KMER);
      fk_table_split_runs(RES.run_table,RES.run_ntable,RES.nruns,TMER_WORD,NTHREADS,beg);
      for (n = 0; n < RES.nruns; n++)
        { const uint8_t *tab = RES.run_table[n];
          const int64   nt = RES.run_ntable[n];
          int64 i = 0;
          for (t = 0; t < NTHREADS; t++)
            { int64 j = i;
              int   f;
              while (j < nt && tab[j*TMER_WORD] < beg[t+1]) j++;
              sprintf(name,"%s/%s.%d.L%d",SORT_PATH,root,n,t);
              f = open(name,O_WRONLY|O_CREAT|O_TRUNC,S_IRWXU|S_IRWXG|S_IRWXO);
              if (f < 0 || (j > i && write(f,tab + i*TMER_WORD,(size_t) (j-i)*TMER_WORD) < 0))
                { fprintf(stderr,"%s: Cannot write to %s.  Enough disk space?\n",Prog_Name,name);
                  Clean_Exit(1);
                }
              close(f);
              i = j;
            }
        }
      free(name); free(beg);
    }
  if (!DO_PROFILE)
    { fkgpu_destroy(CTX); CTX = NULL; }
}

void Merge_Profiles(char *path, char *root)
{ int64_t nr; const int64_t *off; const uint16_t *prof;
  int64_t *rbeg = (int64_t *) malloc(sizeof(int64_t)*(ITHREADS+1));
  int64_t *cnt  = (int64_t *) malloc(sizeof(int64_t)*(ITHREADS+1));
  int t;
  if (VERBOSE)
    fprintf(stderr,"\nPhase 4 (-p option): Writing Profiles\n");
  if (fkgpu_profiles(CTX,&nr,&off,&prof) != 0) fail("fkgpu_profiles");
  if (fkgpu_read_counts(CTX,cnt) != 0) fail("fkgpu_read_counts");
  rbeg[0] = 0;
  for (t = 0; t < ITHREADS; t++) rbeg[t+1] = rbeg[t] + cnt[t];
  if (fk_write_prof(path,root,KMER,ITHREADS,rbeg,off,prof))
    { fprintf(stderr,"%s: Cannot write to %s/%s.prof.  Enough disk space?\n",Prog_Name,path,root);
      Clean_Exit(1);
    }
  free(rbeg); free(cnt);
  fkgpu_destroy(CTX); CTX = NULL;
}
