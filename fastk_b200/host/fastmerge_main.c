/*  fastmerge_main.c -- the Fastmerge command line (Fastmerge.c: merging the k-mer tables and histograms of independent FastK
 *  runs on parts of a data set) over libfastk_gpu.so: the tables are read back into [key][count] records, merged on the GPU
 *  (fkgpu_merge_tables: one key-order sort of all records, then counts of equal k-mers added and saturated) and written with
 *  the same writers as FastK's own table.  Plain C; the profile merge of the reference tool (-p sources) is not part of this.
 *
 *      Fastmerge [-ht] [-T<int(4)>] [-P<dir>] <target> <source>[.hist|.ktab] ...
 *
 *  -h merged histogram, -t merged table (at least one), -T parts of the merged table (Fastmerge.c:26-27,520-540).
 *  Histogram rule (Fastmerge.c:311-331,1009-1027): bin = merged, saturated count; max_inst = the max_inst of every input
 *  histogram + what the unsaturated members of saturated sums stood for.                                             */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <fcntl.h>
#include <unistd.h>
#include "fastk_gpu.h"
#include "fk_files.h"

static char *Prog_Name = "Fastmerge";

static char *strip(const char *a)
{ char *s = strdup(a);
  size_t n = strlen(s);
  if (n > 5 && (strcmp(s+n-5,".hist") == 0 || strcmp(s+n-5,".ktab") == 0 || strcmp(s+n-5,".prof") == 0)) s[n-5] = '\0';
  return s;
}

/* max_inst field of <root>.hist (count.c:1896-1909): int k, int low, int high, int64 ilow, int64 max_inst, ... ; -1 if absent */
static int64_t hist_max_inst(const char *root)
{ char *p = (char *) malloc(strlen(root) + 8);
  int64_t v[2];
  int f, hdr[3];
  sprintf(p,"%s.hist",root);
  f = open(p,O_RDONLY);
  free(p);
  if (f < 0) return -1;
  if (read(f,hdr,sizeof(hdr)) != (ssize_t) sizeof(hdr) || read(f,v,sizeof(v)) != (ssize_t) sizeof(v)) { close(f); return -1; }
  close(f);
  return v[1];
}

int main(int argc, char *argv[])
{ int do_hist = 0, do_table = 0, nthreads = 4, i, narg = 0;
  char *args[4096];
  for (i = 1; i < argc; i++)
    if (argv[i][0] == '-' && argv[i][1] != '\0')
      { char *a = argv[i];
        if (a[1] == 'T') { nthreads = atoi(a+2); if (nthreads <= 0) { fprintf(stderr,"%s: Number of threads must be positive\n",Prog_Name); exit(1); } }
        else if (a[1] == 'P') ;                           /* table cache directory of the reference tool: nothing to cache here */
        else if (a[1] == 'S' || a[1] == '#')
          { fprintf(stderr,"%s: -%c (slices / parts per thread) is not supported by the GPU path\n",Prog_Name,a[1]); exit(1); }
        else
          for (char *c = a+1; *c; c++)
            if (*c == 'h') do_hist = 1;
            else if (*c == 't') do_table = 1;
            else { fprintf(stderr,"%s: -%c is an illegal option\n",Prog_Name,*c); exit(1); }
      }
    else if (narg < 4096) args[narg++] = argv[i];
  if (narg < 3)
    { fprintf(stderr,"\nUsage: %s [-ht] [-T<int(4)>] [-P<dir(/tmp)>] <target> <source>[.hist|.ktab] ...\n\n",Prog_Name);
      fprintf(stderr,"      -h: Produce a merged histogram.\n      -t: Produce a merged k-mer table.\n\n      -T: Use -T threads.\n");
      exit(1);
    }
  if (do_hist + do_table == 0) { fprintf(stderr,"%s: At least one of -h or -t must be set\n",Prog_Name); exit(1); }
  { size_t n = strlen(args[0]);
    if (n > 5 && (strcmp(args[0]+n-5,".hist") == 0 || strcmp(args[0]+n-5,".ktab") == 0 || strcmp(args[0]+n-5,".prof") == 0))
      { fprintf(stderr,"%s: Target name cannot have a .hist, .ktab, or .prof suffix\n",Prog_Name); exit(1); }
  }
  const int ntab = narg - 1;
  uint8_t **tab = (uint8_t **) calloc((size_t) ntab,sizeof(uint8_t *));
  int64_t  *tn  = (int64_t *) calloc((size_t) ntab,sizeof(int64_t));
  int kmer = 0, minval = 0x10000;
  int64_t add_inst = 0;
  for (i = 0; i < ntab; i++)
    { char *src = strip(args[i+1]);
      int k, cut;
      if (fk_read_ktab(src,&k,&cut,&tab[i],&tn[i]))
        { fprintf(stderr,"%s: Cannot open FastK table %s\n",Prog_Name,src); exit(1); }
      if (i == 0) kmer = k;
      else if (k != kmer) { fprintf(stderr,"%s: K-mer tables do not involve the same K\n",Prog_Name); exit(1); }
      if (cut < minval) minval = cut;
      if (do_hist)
        { int64_t mi = hist_max_inst(src);
          if (mi < 0)
            { if (i == 0) { fprintf(stderr,"%s: Warning: no input histograms => overflow count low\n",Prog_Name); do_hist = 2; }
              else if (do_hist == 1) { fprintf(stderr,"%s: Cannot open histogram %s\n",Prog_Name,src); exit(1); }
            }
          else if (do_hist == 1) add_inst += mi;
        }
      free(src);
    }

  fkgpu_config cfg;
  fkgpu_ctx   *ctx;
  fkgpu_result res;
  memset(&cfg,0,sizeof(cfg));
  cfg.kmer = kmer; cfg.do_table = 1; cfg.nthreads = 1;
  cfg.device = getenv("FASTK_GPU") ? atoi(getenv("FASTK_GPU")) : 0;
  if (fkgpu_create(&cfg,&ctx) != 0) { fprintf(stderr,"%s: %s\n",Prog_Name,fkgpu_last_error()); exit(1); }
  if (fkgpu_merge_tables(ctx,(const uint8_t *const *) tab,tn,ntab,do_table,&res) != 0)
    { fprintf(stderr,"%s: %s\n",Prog_Name,fkgpu_last_error()); exit(1); }

  char *target = strdup(args[0]), *dir, *root, *slash = strrchr(target,'/');
  if (slash == NULL) { dir = "."; root = target; } else { *slash = '\0'; dir = target[0] ? target : "/"; root = slash+1; }
  if (do_table && fk_write_ktab(dir,root,kmer,minval,nthreads,res.table,res.ntable))
    { fprintf(stderr,"%s: Cannot write to %s/%s.ktab\n",Prog_Name,dir,root); exit(1); }
  if (do_hist && fk_write_hist(dir,root,kmer,res.hist,res.max_inst + add_inst))
    { fprintf(stderr,"%s: Cannot write to %s/%s.hist\n",Prog_Name,dir,root); exit(1); }
  fkgpu_destroy(ctx);
  for (i = 0; i < ntab; i++) free(tab[i]);
  free(tab); free(tn); free(target);
  return 0;
}
