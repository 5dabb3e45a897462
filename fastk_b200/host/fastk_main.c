/*  fastk_main.c -- the FastK command line over libfastk_gpu (C host, no CUDA in this file).
 *
 *    FastK [-k<int(40)>] [-t[<int(1)>]] [-p] [-c] [-bc<int>] [-v] [-N<path_name>] [-P<dir>] [-M<int>] [-T<int(4)>]
 *          <source>[.fasta|.fastq|.fa|.fq][.gz] ...
 *
 *  Same option grammar and output names as the reference driver (FastK.c:223-357): <root>.hist always,
 *  <root>.ktab + hidden parts with -t, <root>.prof + hidden parts with -p.  The stages it sequences are the
 *  reference's (FastK.c:498-540) with the replaced ones behind the C ABI:
 *      Split_Kmers / Distribute_Block -> fkgpu_ingest      Sorting -> fkgpu_finish
 *      Merge_Tables                   -> fk_write_ktab     Merge_Profiles -> fkgpu_profiles + fk_write_prof
 *  Input is read by ITHREADS host threads over byte ranges of the files (the role of io.c:2280-2600,574-759);
 *  FASTA / FASTQ, optionally gzip'd (one thread per .gz file).  -p:<table> (relative profiles), BAM/CRAM and
 *  Dazzler inputs are not handled here -- link the reference's own io.c with fastk_shim.c for those
 *  (INTEGRATION.md).  Errors: message on stderr, partial outputs removed, exit 1 (Clean_Exit, FastK.c:181-221).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <pthread.h>
#include <unistd.h>
#include <fcntl.h>
#include <libgen.h>
#include <time.h>
#include <sys/stat.h>
#include <sys/resource.h>
#include <zlib.h>

#include "fastk_gpu.h"
#include "fk_files.h"

static char *Prog_Name = "FastK";

static int   VERBOSE, COMPRESS, KMER = 40, DO_TABLE, DO_PROFILE, BC_PREFIX, NTHREADS = 4, ITHREADS;
static char   *PRO_NAME;               /* -p:<table>: profiles relative to this k-mer table */
static int64_t SORT_MEMORY;            /* -M in bytes; 0 = not given: use what the device has */
static char *OUT_NAME;
static char  OUT_DIR[4096], OUT_ROOT[4096];
static int   have_out;

/*  FASTK_GPUS=<n> counts on n GPUs (devices FASTK_GPU .. FASTK_GPU+n-1): one context per GPU, joined by one NCCL communicator
    inside the library; reader thread t feeds context t mod n.  Minimizer buckets and key ranges are exchanged over NVLink
    and every GPU returns its key range of the table: the rank-ordered ranges are the runs the table writer merges.       */
#define MAX_GPUS 16
static fkgpu_ctx *CTXS[MAX_GPUS];
static int        NGPUS = 1;
#define CTX (CTXS[0])

static void Clean_Exit(int status)
{ int g;
  if (have_out) fk_remove_outputs(OUT_DIR,OUT_ROOT);
  if (NGPUS > 1) _exit(status);           /* peers may sit in a collective: do not wait for their contexts */
  for (g = 0; g < NGPUS; g++)
    if (CTXS[g]) fkgpu_destroy(CTXS[g]);
  exit(status);
}

typedef struct { int g; const uint8_t *id; int fetch; fkgpu_result res; int rc; char err[1024]; } Gpu_Job;

static void *comm_thread(void *arg)
{ Gpu_Job *J = (Gpu_Job *) arg;
  J->rc = fkgpu_comm_init(CTXS[J->g],NGPUS,J->g,J->id);
  if (J->rc) snprintf(J->err,sizeof(J->err),"%s",fkgpu_last_error());
  return NULL;
}

static void *finish_thread(void *arg)
{ Gpu_Job *J = (Gpu_Job *) arg;
  J->rc = fkgpu_finish(CTXS[J->g],J->fetch,&J->res);
  if (J->rc) snprintf(J->err,sizeof(J->err),"%s",fkgpu_last_error());
  return NULL;
}

static double now(void)
{ struct timespec t;
  clock_gettime(CLOCK_MONOTONIC,&t);
  return t.tv_sec + 1e-9*t.tv_nsec;
}

/* ---- input ------------------------------------------------------------------------------------ */

#define DT_BLOCK 1000000          /* io.c:64-66 */
#define DT_MINIM  100000
#define DT_READS   10000

typedef struct
  { char  *path;
    int64_t size;
    int    zipd, fastq;
  } File_Object;

typedef struct
  { int          tid;
    File_Object *files;
    int          bfile, efile;     /* files [bfile, efile] */
    int64_t      bpos, epos;       /* byte range in the first / last file */
    int64_t      nreads, nbases;
    int          err;
  } Reader;

typedef struct
  { char   *bases;
    int32_t *boff;
    int      nreads;
    int64_t  fill;
  } Block;

static void block_flush(Reader *R, Block *B, int rem)
{ if (B->nreads > 0)
    { if (fkgpu_ingest(CTXS[R->tid % NGPUS],R->tid / NGPUS,B->bases,B->boff,B->nreads,rem) != 0)
        { fprintf(stderr,"%s: %s\n",Prog_Name,fkgpu_last_error());
          R->err = 1;
        }
    }
  B->nreads = 0;
  B->fill = 0;
  B->boff[0] = 0;
}

/* append one whole read; long reads are cut into <= DT_BLOCK pieces that overlap by k-1 (io.c:296-333) */
static void block_add(Reader *R, Block *B, char *seq, int64_t len)
{ if (COMPRESS && len > 0)                       /* io.c:284-294 */
    { int64_t i, n = 1;
      char x = seq[0];
      for (i = 1; i < len; i++)
        if (seq[i] != x) seq[n++] = x = seq[i];
      len = n;
    }
  R->nreads += 1;
  R->nbases += len;
  for (;;)
    { int64_t room = DT_BLOCK - B->fill - 1;
      if (B->nreads >= DT_READS || (len > room && room < DT_MINIM))
        { block_flush(R,B,0);
          room = DT_BLOCK - 1;
        }
      if (len <= room)
        { memcpy(B->bases + B->fill,seq,len);
          B->fill += len;
          B->bases[B->fill++] = '\0';
          B->boff[++B->nreads] = (int32_t) B->fill;
          return;
        }
      /* piece of a long read: room bases now, the rest continues with a k-1 overlap */
      memcpy(B->bases + B->fill,seq,room);
      B->fill += room;
      B->bases[B->fill++] = '\0';
      B->boff[++B->nreads] = (int32_t) B->fill;
      block_flush(R,B,1);
      seq += room - (KMER-1);
      len -= room - (KMER-1);
    }
}

/* the io.c:678-734 automaton over an arbitrary byte source */
typedef struct
  { int   state, fastq;
    char *seq; int64_t slen, smax;
  } Parser;
enum { QAT, HSKP, QSEQ, QPLS, QSKP, AEOL, ASEQ };

static inline void seq_append(Parser *P, const char *s, int64_t n)
{ if (P->slen + n > P->smax)
    { P->smax = 2*(P->slen + n) + 65536;
      P->seq = (char *) realloc(P->seq,P->smax);
      if (P->seq == NULL) { fprintf(stderr,"%s: Out of memory (sequence buffer of %lld bytes)\n",Prog_Name,(long long) P->smax); exit(1); }
    }
  memcpy(P->seq + P->slen,s,(size_t) n);
  P->slen += n;
}

/* Same transitions as the byte-at-a-time automaton of io.c:678-734, taken a line segment at a time: header and
   quality lines are skipped with memchr, sequence lines are appended with one memcpy each.                      */
static void parse_bytes(Reader *R, Block *B, Parser *P, const char *buf, int64_t n)
{ const char *p = buf, *e = buf + n, *q;
  while (p < e)
    switch (P->state)
    { case QAT:  P->state = HSKP; p++; break;                      /* the '@' / '>' that opens a record */
      case HSKP: q = (const char *) memchr(p,'\n',(size_t) (e-p));
                 if (q == NULL) { p = e; break; }
                 P->state = P->fastq ? QSEQ : ASEQ; p = q+1;
                 break;
      case QSEQ: q = (const char *) memchr(p,'\n',(size_t) (e-p));
                 seq_append(P,p,(q ? q : e) - p);
                 if (q == NULL) { p = e; break; }
                 block_add(R,B,P->seq,P->slen); P->slen = 0; P->state = QPLS; p = q+1;
                 break;
      case QPLS: q = (const char *) memchr(p,'\n',(size_t) (e-p));
                 if (q == NULL) { p = e; break; }
                 P->state = QSKP; p = q+1;
                 break;
      case QSKP: q = (const char *) memchr(p,'\n',(size_t) (e-p));
                 if (q == NULL) { p = e; break; }
                 P->state = QAT; p = q+1;
                 break;
      case AEOL: if (*p == '>') { block_add(R,B,P->seq,P->slen); P->slen = 0; P->state = HSKP; p++; }
                 else if (*p == '\n') p++;
                 else P->state = ASEQ;                              /* first base of the next sequence line: ASEQ takes it */
                 break;
      case ASEQ: q = (const char *) memchr(p,'\n',(size_t) (e-p));
                 seq_append(P,p,(q ? q : e) - p);
                 if (q == NULL) { p = e; break; }
                 P->state = AEOL; p = q+1;
                 break;
    }
}

static void *reader_thread(void *arg)
{ Reader *R = (Reader *) arg;
  Block   B;
  Parser  P;
  char   *buf = (char *) malloc(1 << 22);
  int     f;

  B.bases = (char *) malloc(DT_BLOCK + 8);
  B.boff  = (int32_t *) malloc(sizeof(int32_t)*(DT_READS+2));
  memset(&P,0,sizeof(P));
  if (buf == NULL || B.bases == NULL || B.boff == NULL)
    { fprintf(stderr,"%s: Out of memory (reader thread buffers)\n",Prog_Name);
      R->err = 1;
      free(buf); free(B.bases); free(B.boff);
      return NULL;
    }
  B.nreads = 0; B.fill = 0; B.boff[0] = 0;
  for (f = R->bfile; f <= R->efile && !R->err; f++)
    { File_Object *F = R->files + f;
      int64_t beg = (f == R->bfile) ? R->bpos : 0;
      int64_t end = (f == R->efile) ? R->epos : F->size;
      P.state = QAT; P.fastq = F->fastq; P.slen = 0;
      if (F->zipd)
        { gzFile g = gzopen(F->path,"rb");
          int n;
          if (g == NULL) { fprintf(stderr,"%s: Cannot open %s\n",Prog_Name,F->path); R->err = 1; break; }
          while ((n = gzread(g,buf,1 << 22)) > 0) parse_bytes(R,&B,&P,buf,n);
          { int e = 0;                       /* corrupt or truncated .gz: a partial count must not look like a result */
            const char *msg = gzerror(g,&e);
            if (n < 0 || !gzeof(g) || (e != Z_OK && e != Z_STREAM_END))
              { fprintf(stderr,"%s: Error reading %s: %s\n",Prog_Name,F->path,(e != Z_OK && msg != NULL && msg[0]) ? msg : "truncated compressed file");
                R->err = 1;
              }
          }
          gzclose(g);
        }
      else
        { int fd = open(F->path,O_RDONLY);
          int64_t pos = beg;
          if (fd < 0) { fprintf(stderr,"%s: Cannot open %s\n",Prog_Name,F->path); R->err = 1; break; }
          lseek(fd,beg,SEEK_SET);
          while (pos < end)
            { int64_t want = end-pos; ssize_t n;
              if (want > (1 << 22)) want = 1 << 22;
              n = read(fd,buf,(size_t) want);
              if (n <= 0)
                { fprintf(stderr,"%s: Error reading %s (%s at byte %lld of %lld)\n",Prog_Name,F->path,
                          n < 0 ? "read failed" : "file is shorter than it was",(long long) pos,(long long) end);
                  R->err = 1;
                  break;
                }
              parse_bytes(R,&B,&P,buf,n);
              pos += n;
            }
          close(fd);
        }
      if (P.state == AEOL)                         /* io.c:737-738 */
        { block_add(R,&B,P.seq,P.slen); P.slen = 0; }
    }
  block_flush(R,&B,0);
  free(buf); free(B.bases); free(B.boff); free(P.seq);
  return NULL;
}

/* first record start at or after byte `pos` of an uncompressed file (role of fast_nearest, io.c:409-470) */
static int64_t record_start(File_Object *F, int64_t pos)
{ FILE *fp;
  char *line = NULL; size_t cap = 0; ssize_t n;
  int64_t at;
  if (pos <= 0) return 0;
  if (pos >= F->size) return F->size;
  fp = fopen(F->path,"rb");
  if (fp == NULL) return F->size;
  fseeko(fp,pos-1,SEEK_SET);
  if (fgetc(fp) != '\n')                           /* move to the next line start */
    { if ((n = getline(&line,&cap,fp)) < 0) { fclose(fp); free(line); return F->size; } }
  at = ftello(fp);
  if (!F->fastq)
    { while ((n = getline(&line,&cap,fp)) >= 0)
        { if (line[0] == '>') break;
          at += n;
        }
      if (n < 0) at = F->size;
    }
  else
    { /* a record starts at a line beginning with '@' whose line after next begins with '+' */
      char *l[4] = { NULL, NULL, NULL, NULL }; size_t c[4] = { 0,0,0,0 }; ssize_t m[4];
      int64_t p = at;
      for (;;)
        { int i, ok = 1;
          fseeko(fp,p,SEEK_SET);
          for (i = 0; i < 3; i++)
            if ((m[i] = getline(&l[i],&c[i],fp)) < 0) { ok = 0; break; }
          if (!ok) { at = F->size; break; }
          if (l[0][0] == '@' && l[2][0] == '+') { at = p; break; }
          p += m[0];
        }
      for (int i = 0; i < 4; i++) free(l[i]);
    }
  fclose(fp); free(line);
  return at;
}

/* ---- main ------------------------------------------------------------------------------------- */

static void strip_suffix(char *root)
{ static const char *sfx[] = { ".fastq.gz", ".fasta.gz", ".fq.gz", ".fa.gz", ".fastq", ".fasta", ".fq", ".fa", NULL };
  size_t L = strlen(root);
  for (int i = 0; sfx[i]; i++)
    { size_t S = strlen(sfx[i]);
      if (L > S && strcmp(root+L-S,sfx[i]) == 0) { root[L-S] = '\0'; return; }
    }
}

static int find_file(const char *arg, File_Object *F)
{ static const char *sfx[] = { "", ".fasta", ".fastq", ".fa", ".fq", ".fasta.gz", ".fastq.gz", ".fa.gz", ".fq.gz", NULL };
  char path[4096];
  struct stat st;
  for (int i = 0; sfx[i]; i++)
    { snprintf(path,sizeof(path),"%s%s",arg,sfx[i]);
      if (stat(path,&st) == 0 && S_ISREG(st.st_mode))
        { size_t L = strlen(path);
          gzFile g;
          int c;
          F->path = strdup(path);
          F->size = st.st_size;
          F->zipd = (L > 3 && strcmp(path+L-3,".gz") == 0);
          g = gzopen(path,"rb");
          c = g ? gzgetc(g) : -1;
          if (g) gzclose(g);
          if (c != '>' && c != '@' && st.st_size > 0)
            { fprintf(stderr,"%s: %s is neither fasta nor fastq (BAM/CRAM/Dazzler need the reference's io.c)\n",Prog_Name,path);
              return 1;
            }
          F->fastq = (c == '@');
          return 0;
        }
    }
  fprintf(stderr,"%s: Cannot find %s as a fasta or fastq file\n",Prog_Name,arg);
  return 1;
}

int main(int argc, char *argv[])
{ int   i, nfiles = 0;
  char *files[4096];
  char *spath = NULL;
  double t0 = now(), tc, t1, t2, t3, t4;

  for (i = 1; i < argc; i++)
    if (argv[i][0] == '-' && argv[i][1] != '\0')
      { char *a = argv[i];
        switch (a[1])
        { case 'k': KMER = atoi(a+2); if (KMER <= 0) { fprintf(stderr,"%s: K-mer length must be positive\n",Prog_Name); exit(1); } break;
          case 'T': NTHREADS = atoi(a+2); if (NTHREADS <= 0) { fprintf(stderr,"%s: Number of threads must be positive\n",Prog_Name); exit(1); } break;
          case 'M': SORT_MEMORY = (int64_t) (atof(a+2) * 1073741824.);    /* SORT_MEMORY (FastK.c:353-361), here: device memory for the count's working buffers; an input that needs more is counted in rounds */
                    if (SORT_MEMORY <= 0) { fprintf(stderr,"%s: Memory must be positive\n",Prog_Name); exit(1); } break;
          case 'P': spath = a+2; break;
          case 'N': OUT_NAME = a+2; break;
          case 'b':
            if (a[2] != 'c') { fprintf(stderr,"\n%s: -%s is not a legal optional argument\n",Prog_Name,a); exit(1); }
            BC_PREFIX = atoi(a+3);
            break;
          case 'p':
            if (a[2] == ':') { PRO_NAME = a+3; DO_PROFILE = 1; break; }      /* relative profiles (FastK.c:269-281) */
            /* fall through */
          default:
            if (a[1] == 't' && isdigit((unsigned char) a[2])) { DO_TABLE = atoi(a+2); break; }
            for (char *c = a+1; *c; c++)
              switch (*c)
              { case 'v': VERBOSE = 1; break;
                case 'c': COMPRESS = 1; break;
                case 'p': DO_PROFILE = 1; break;
                case 't': if (DO_TABLE == 0) DO_TABLE = 1; break;
                default: fprintf(stderr,"\n%s: -%c is an illegal option\n",Prog_Name,*c); exit(1);
              }
        }
      }
    else if (nfiles < 4096)
      files[nfiles++] = argv[i];

  if (nfiles == 0)
    { fprintf(stderr,"\nUsage: %s [-k<int(40)>] [-t[<int(1)>]] [-p] [-c] [-bc<int>]\n",Prog_Name);
      fprintf(stderr,"             [-v] [-N<path_name>] [-P<dir>] [-M<int>] [-T<int(4)>]\n");
      fprintf(stderr,"               <source>[.fasta|.fastq][.gz] ...\n\n");
      fprintf(stderr,"      -v: Verbose mode, output statistics as proceed.\n");
      fprintf(stderr,"      -T: Use -T threads.\n");
      fprintf(stderr,"      -N: Use given path for output directory and root name prefix.\n");
      fprintf(stderr,"      -P: (accepted; the GPU path keeps no temporary files)\n");
      fprintf(stderr,"      -M: (accepted; the device arena is sized from the input)\n");
      fprintf(stderr,"\n      -k: k-mer size.\n");
      fprintf(stderr,"      -t: Produce table of sorted k-mers & counts >= level specified\n");
      fprintf(stderr,"      -p: Produce sequence count profiles\n");
      fprintf(stderr,"     -bc: Ignore prefix of each read of given length (e.g. bar code)\n");
      fprintf(stderr,"      -c: Homopolymer compress every sequence\n");
      exit(1);
    }
  if (spath != NULL)
    { struct stat st;
      if (stat(spath,&st) != 0 || !S_ISDIR(st.st_mode))
        { fprintf(stderr,"\n%s: -P option: cannot open directory %s\n",Prog_Name,spath); exit(1); }
    }

  File_Object *F = (File_Object *) calloc(nfiles,sizeof(File_Object));
  int64_t work = 0;
  int     anyzip = 0;
  for (i = 0; i < nfiles; i++)
    { if (find_file(files[i],F+i)) exit(1);
      if (F[i].fastq != F[0].fastq) { fprintf(stderr,"%s: All files must be of the same type\n",Prog_Name); exit(1); }
      work += F[i].size;
      anyzip |= F[i].zipd;
    }
  { char *src = strdup(OUT_NAME ? OUT_NAME : F[0].path), *s2 = strdup(src);
    snprintf(OUT_DIR,sizeof(OUT_DIR),"%s",dirname(src));
    snprintf(OUT_ROOT,sizeof(OUT_ROOT),"%s",basename(s2));
    if (!OUT_NAME) strip_suffix(OUT_ROOT);
    free(src); free(s2);
  }

  /* input threads: io.c:2373-2378 (gz: whole files per thread), io.c:2424-2430 (tiny inputs use fewer) */
  if (anyzip) ITHREADS = nfiles < NTHREADS ? nfiles : NTHREADS;
  else
    { ITHREADS = NTHREADS;
      if (work/NTHREADS < 200000) { ITHREADS = (int) (work/200000); if (ITHREADS <= 0) ITHREADS = 1; }
    }

  fkgpu_config cfg;
  memset(&cfg,0,sizeof(cfg));
  cfg.kmer = KMER; cfg.do_table = DO_TABLE; cfg.do_profile = DO_PROFILE; cfg.bc_prefix = BC_PREFIX;
  cfg.device = getenv("FASTK_GPU") ? atoi(getenv("FASTK_GPU")) : 0;
  NGPUS = getenv("FASTK_GPUS") ? atoi(getenv("FASTK_GPUS")) : 1;
  if (NGPUS < 1 || NGPUS > MAX_GPUS) { fprintf(stderr,"%s: FASTK_GPUS must be in [1,%d]\n",Prog_Name,MAX_GPUS); exit(1); }
  if (NGPUS > ITHREADS) NGPUS = ITHREADS;
  if (NGPUS > 1 && DO_PROFILE) { fprintf(stderr,"%s: -p needs the whole table on one GPU: run without FASTK_GPUS\n",Prog_Name); exit(1); }
  cfg.nthreads = (ITHREADS + NGPUS - 1) / NGPUS; cfg.reserve_bases = work / NGPUS + (NGPUS > 1 ? work / (8*NGPUS) : 0); cfg.mem_limit = SORT_MEMORY;
  { int g;
    for (g = 0; g < NGPUS; g++)
      { fkgpu_config cg = cfg;
        cg.device = cfg.device + g;
        if (fkgpu_create(&cg,&CTXS[g]) != 0)
          { fprintf(stderr,"%s: %s\n",Prog_Name,fkgpu_last_error()); exit(1); }
      }
    if (NGPUS > 1)
      { uint8_t   id[FKGPU_COMM_ID_BYTES];
        Gpu_Job   job[MAX_GPUS];
        pthread_t th[MAX_GPUS];
        if (fkgpu_comm_id(id) != 0) { fprintf(stderr,"%s: %s\n",Prog_Name,fkgpu_last_error()); exit(1); }
        for (g = 0; g < NGPUS; g++) { job[g].g = g; job[g].id = id; pthread_create(th+g,NULL,comm_thread,job+g); }
        for (g = 0; g < NGPUS; g++) pthread_join(th[g],NULL);
        for (g = 0; g < NGPUS; g++)
          if (job[g].rc) { fprintf(stderr,"%s: GPU %d: %s\n",Prog_Name,g,job[g].err); exit(1); }
        if (VERBOSE) fprintf(stderr,"  Counting on %d GPUs (one NCCL communicator inside the library)\n",NGPUS);
      }
  }
  if (PRO_NAME != NULL)
    { /* what Split_Table does for the reference (split.c:1943-2131): bring the table to where the profiles are made */
      int tk, tcut; uint8_t *trec; int64_t tn;
      if (fk_read_ktab(PRO_NAME,&tk,&tcut,&trec,&tn))
        { fprintf(stderr,"%s: Cannot open FastK table %s\n",Prog_Name,PRO_NAME); exit(1); }
      if (tk != KMER)
        { fprintf(stderr,"%s: -p table k-mer size (%d) != k-mer specified (%d)\n",Prog_Name,tk,KMER); exit(1); }
      if (DO_TABLE && VERBOSE) fprintf(stderr,"%s: Warning: -p:%s overides -t option\n",Prog_Name,PRO_NAME);
      DO_TABLE = 0;
      if (fkgpu_load_profile_table(CTX,trec,tn) != 0)
        { fprintf(stderr,"%s: %s\n",Prog_Name,fkgpu_last_error()); exit(1); }
      free(trec);
      if (VERBOSE) fprintf(stderr,"  Profiles relative to %s: %lld %d-mers with counts >= %d\n",PRO_NAME,(long long) tn,tk,tcut);
    }
  have_out = 1;
  fk_remove_outputs(OUT_DIR,OUT_ROOT);
  tc = now();

  if (VERBOSE)
    fprintf(stderr,"\nPhase 1: Reading %d file(s) with %d thread(s) into the GPU k-mer counter\n",nfiles,ITHREADS);

  Reader   *R = (Reader *) calloc(ITHREADS,sizeof(Reader));
  pthread_t th[ITHREADS];
  if (anyzip)
    for (i = 0; i < ITHREADS; i++)
      { R[i].bfile = (i*nfiles)/ITHREADS; R[i].efile = ((i+1)*nfiles)/ITHREADS - 1;
        R[i].bpos = 0; R[i].epos = F[R[i].efile].size;
      }
  else
    { /* cut the concatenation of all files into ITHREADS byte ranges at record starts */
      int64_t cut[ITHREADS+1]; int cf[ITHREADS+1];
      cut[0] = 0; cf[0] = 0;
      for (i = 1; i < ITHREADS; i++)
        { int64_t w = (work*i)/ITHREADS; int f = 0;
          while (f < nfiles-1 && w >= F[f].size) { w -= F[f].size; f++; }
          cf[i] = f; cut[i] = record_start(F+f,w);
        }
      cf[ITHREADS] = nfiles-1; cut[ITHREADS] = F[nfiles-1].size;
      for (i = 0; i < ITHREADS; i++)
        { R[i].bfile = cf[i]; R[i].bpos = cut[i]; R[i].efile = cf[i+1]; R[i].epos = cut[i+1];
          if (R[i].efile > R[i].bfile && R[i].epos == 0) { R[i].efile -= 1; R[i].epos = F[R[i].efile].size; }
        }
    }
  for (i = 0; i < ITHREADS; i++)
    { R[i].tid = i; R[i].files = F;
      pthread_create(th+i,NULL,reader_thread,R+i);
    }
  int bad = 0;
  int64_t nreads = 0, nbases = 0;
  for (i = 0; i < ITHREADS; i++)
    { pthread_join(th[i],NULL);
      bad |= R[i].err; nreads += R[i].nreads; nbases += R[i].nbases;
    }
  if (bad) Clean_Exit(1);
  t1 = now();
  if (VERBOSE)
    fprintf(stderr,"  %lld reads, %lld bases in %.3fs\n",(long long) nreads,(long long) nbases,t1-t0);

  if (VERBOSE) fprintf(stderr,"\nPhase 2: Sorting & Counting K-mers on the GPU\n");
  fkgpu_result res;
  const uint8_t *mruns[MAX_GPUS]; int64_t mrun_n[MAX_GPUS];
  if (NGPUS == 1)
    { if (fkgpu_finish(CTX,DO_TABLE > 0,&res) != 0)
        { fprintf(stderr,"%s: %s\n",Prog_Name,fkgpu_last_error()); Clean_Exit(1); }
    }
  else
    { /* the finish is collective: one thread per GPU; GPU g returns the g-th key range of the table */
      static Gpu_Job job[MAX_GPUS];
      pthread_t th[MAX_GPUS];
      int g;
      for (g = 0; g < NGPUS; g++) { job[g].g = g; job[g].fetch = DO_TABLE > 0; pthread_create(th+g,NULL,finish_thread,job+g); }
      for (g = 0; g < NGPUS; g++) pthread_join(th[g],NULL);
      for (g = 0; g < NGPUS; g++)
        if (job[g].rc) { fprintf(stderr,"%s: GPU %d: %s\n",Prog_Name,g,job[g].err); Clean_Exit(1); }
      res = job[0].res;
      res.ntable = 0; res.nbases = 0; res.nreads = 0;
      for (g = 0; g < NGPUS; g++)
        { mruns[g] = job[g].res.table; mrun_n[g] = job[g].res.ntable;
          res.ntable += job[g].res.ntable; res.nbases += job[g].res.nbases; res.nreads += job[g].res.nreads;
          if (job[g].res.ms_total > res.ms_total) res.ms_total = job[g].res.ms_total;
        }
      res.nruns = NGPUS; res.run_table = mruns; res.run_ntable = mrun_n;
    }
  t2 = now();
  if (VERBOSE)
    { fprintf(stderr,"  %lld %d-mers, %lld distinct; device %.3f ms (pack %.3f ms), wall %.3fs\n",
              (long long) res.nkmers,KMER,(long long) res.ndistinct,res.ms_total,res.ms_pack,t2-t1);
      fprintf(stderr,"  %.3f Gbases/s on the device\n",res.ms_total > 0 ? res.nbases/1e6/res.ms_total : 0.);
      if (res.nruns > 1 && NGPUS == 1) fprintf(stderr,"  Counted in %d rounds (sorted runs merged while the table parts are written)\n",res.nruns);
    }

  if (PRO_NAME == NULL && fk_write_hist(OUT_DIR,OUT_ROOT,KMER,res.hist,res.max_inst))
    { fprintf(stderr,"%s: Cannot write to %s/%s.hist.  Enough disk space?\n",Prog_Name,OUT_DIR,OUT_ROOT); Clean_Exit(1); }
  if (DO_TABLE > 0)
    { if (VERBOSE) fprintf(stderr,"\nPhase 3 (-t option): Writing K-mer Table Parts\n");
      if (fk_write_ktab_runs(OUT_DIR,OUT_ROOT,KMER,DO_TABLE,NTHREADS,res.run_table,res.run_ntable,res.nruns))
        { fprintf(stderr,"%s: Cannot write to %s/%s.ktab.  Enough disk space?\n",Prog_Name,OUT_DIR,OUT_ROOT); Clean_Exit(1); }
      if (VERBOSE)
        fprintf(stderr,"  There are %lld %d-mers that occur %d-or-more times\n",(long long) res.ntable,KMER,DO_TABLE);
    }
  if (DO_PROFILE)
    { int64_t nr; const int64_t *off; const uint16_t *prof;
      int64_t rbeg[ITHREADS+1];
      if (VERBOSE) fprintf(stderr,"\nPhase 4 (-p option): Writing Profiles\n");
      if (fkgpu_profiles(CTX,&nr,&off,&prof) != 0)
        { fprintf(stderr,"%s: %s\n",Prog_Name,fkgpu_last_error()); Clean_Exit(1); }
      rbeg[0] = 0;
      for (i = 0; i < ITHREADS; i++) rbeg[i+1] = rbeg[i] + R[i].nreads;
      if (rbeg[ITHREADS] != nr)
        { fprintf(stderr,"%s: internal: profile read count %lld != %lld\n",Prog_Name,(long long) nr,(long long) rbeg[ITHREADS]); Clean_Exit(1); }
      if (fk_write_prof(OUT_DIR,OUT_ROOT,KMER,ITHREADS,rbeg,off,prof))
        { fprintf(stderr,"%s: Cannot write to %s/%s.prof.  Enough disk space?\n",Prog_Name,OUT_DIR,OUT_ROOT); Clean_Exit(1); }
    }
  t3 = now();
  have_out = 0;
  { int g; for (g = 0; g < NGPUS; g++) fkgpu_destroy(CTXS[g]); }
  t4 = now();
  if (VERBOSE)
    { struct rusage ru;
      getrusage(RUSAGE_SELF,&ru);
      fprintf(stderr,"\nTotal Resources:  %.3fs wall (init %.3f, read %.3f, count %.3f, write %.3f, release %.3f)  %ldMB\n",
              t4-t0,tc-t0,t1-tc,t2-t1,t3-t2,t4-t3,ru.ru_maxrss/1024);
    }
  return 0;
}
