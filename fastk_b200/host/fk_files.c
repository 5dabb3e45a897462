/*  fk_files.c -- see fk_files.h.  Plain C, no CUDA. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include <pthread.h>
#include <dirent.h>
#include "fk_files.h"

static int put(int f, const void *buf, int64_t n)
{ const char *p = (const char *) buf;
  while (n > 0)
    { ssize_t w = write(f,p,(size_t) (n > (1 << 30) ? (1 << 30) : n));
      if (w <= 0) return 1;
      p += w; n -= w;
    }
  return 0;
}

int fk_idx_bytes(int64_t nentries, int kmer)
{ if (nentries > 0x4000000ll && kmer >= 12) return 3;
  if (nentries >= 0x40000ll && kmer >= 8) return 2;
  return 1;
}

/* index of the first entry whose first key byte is >= x (entries are sorted) */
static int64_t first_byte_lower_bound(const uint8_t *entries, int64_t n, int tw, int x)
{ int64_t lo = 0, hi = n;
  while (lo < hi)
    { int64_t mid = (lo+hi) >> 1;
      if (entries[mid*tw] < x) lo = mid+1; else hi = mid;
    }
  return lo;
}

void fk_table_split_runs(const uint8_t *const *runs, const int64_t *ns, int nruns, int tw, int nparts, int *beg)
{ int64_t asize = 0, sum = 0, thr;
  int64_t *prev = (int64_t *) calloc((size_t) (nruns > 0 ? nruns : 1),sizeof(int64_t));
  int     x, r, m = 0;
  for (r = 0; r < nruns; r++) asize += ns[r] * tw;
  thr = asize / nparts;
  beg[0] = 0;
  for (x = 0; x < 256; x++)
    { for (r = 0; r < nruns; r++)
        { int64_t next = first_byte_lower_bound(runs[r],ns[r],tw,x+1);       /* part[x] = sum over the runs of (next - prev) * tw */
          sum += (next - (prev ? prev[r] : 0)) * tw;
          if (prev) prev[r] = next;
        }
      if (sum >= thr && m < nparts)
        { beg[++m] = x+1;
          thr = (asize * (m+1)) / nparts;
        }
    }
  while (m < nparts) beg[++m] = 256;
  beg[nparts] = 256;
  free(prev);
}

void fk_table_split(const uint8_t *entries, int64_t n, int tw, int nparts, int *beg)
{ fk_table_split_runs(&entries,&n,1,tw,nparts,beg); }

int fk_write_hist(const char *dir, const char *root, int kmer, const int64_t *hist, int64_t max_inst)
{ char name[4096];
  int  f, v, bad = 0;
  snprintf(name,sizeof(name),"%s/%s.hist",dir,root);
  f = open(name,O_WRONLY|O_CREAT|O_TRUNC,0755);
  if (f < 0) return 1;
  bad |= put(f,&kmer,sizeof(int));
  v = 1;      bad |= put(f,&v,sizeof(int));
  v = 0x7fff; bad |= put(f,&v,sizeof(int));
  bad |= put(f,hist+1,sizeof(int64_t));
  bad |= put(f,&max_inst,sizeof(int64_t));
  bad |= put(f,hist+1,0x7fff*sizeof(int64_t));
  close(f);
  return bad;
}

/* one hidden table part per thread: parts are separate files, and because they are cut on first-byte boundaries no
   prefix-index slot is shared between two parts (table.c:257 relies on the same fact).  The table may arrive as several
   sorted runs with disjoint keys (a multi-round count): the part's slices of the runs are merged on the fly, the role of
   table.c:240-313 for the NPARTS part files.                                                                        */
typedef struct
  { const char *dir, *root; const uint8_t *const *runs; int nruns;
    int64_t *i, *j; int64_t *pindex;
    int kmer, tw, ib, t, bad;
  } Ktab_Job;

static void *ktab_part_thread(void *arg)
{ Ktab_Job *J = (Ktab_Job *) arg;
  const int pw = J->tw - J->ib, kb = J->tw - 2;
  const size_t cap = 1 << 22;
  size_t   fill = 0;
  uint8_t *buf = (uint8_t *) malloc(cap + 64);
  char     name[4096];
  int64_t  m = 0;
  int      f, r, live = 0;
  for (r = 0; r < J->nruns; r++) { m += J->j[r] - J->i[r]; live += (J->j[r] > J->i[r]); }
  snprintf(name,sizeof(name),"%s/.%s.ktab.%d",J->dir,J->root,J->t+1);
  f = open(name,O_WRONLY|O_CREAT|O_TRUNC,0700);
  if (f < 0 || buf == NULL) { J->bad = 1; free(buf); if (f >= 0) close(f); return NULL; }
  J->bad |= put(f,&J->kmer,sizeof(int));
  J->bad |= put(f,&m,sizeof(int64_t));
  while (live > 0)
    { const uint8_t *e = NULL;
      int64_t idx = 0, stop;
      int b, best = -1;
      /* the run whose next entry is smallest; with one live run its whole remainder goes out without comparisons */
      for (r = 0; r < J->nruns; r++)
        if (J->i[r] < J->j[r])
          { const uint8_t *x = J->runs[r] + J->i[r]*J->tw;
            if (best < 0 || memcmp(x,e,(size_t) kb) < 0) { best = r; e = x; }
          }
      stop = (live == 1) ? J->j[best] : J->i[best] + 1;
      for ( ; J->i[best] < stop; J->i[best]++)
        { e = J->runs[best] + J->i[best]*J->tw;
          idx = 0;
          for (b = 0; b < J->ib; b++) idx = (idx << 8) | e[b];
          J->pindex[idx] += 1;
          memcpy(buf+fill,e+J->ib,pw);
          fill += pw;
          if (fill + pw > cap) { J->bad |= put(f,buf,fill); fill = 0; }
        }
      if (J->i[best] >= J->j[best]) live--;
    }
  J->bad |= put(f,buf,fill);
  free(buf);
  close(f);
  return NULL;
}

int fk_write_ktab_runs(const char *dir, const char *root, int kmer, int cutoff, int nparts,
                       const uint8_t *const *runs, const int64_t *ns, int nruns)
{ const int kb = (2*kmer+7) >> 3, tw = kb+2;
  int64_t  n = 0, *pindex, *cur;
  int      ib, r;
  int      *beg = (int *) malloc(sizeof(int)*(nparts+1));
  Ktab_Job *job = (Ktab_Job *) calloc((size_t) nparts,sizeof(Ktab_Job));
  pthread_t *th = (pthread_t *) malloc(sizeof(pthread_t)*(size_t) nparts);
  char     name[4096];
  int64_t  x, ilen;
  int      f, t, bad = 0;

  for (r = 0; r < nruns; r++) n += ns[r];
  ib = fk_idx_bytes(n,kmer);
  ilen = 1ll << (8*ib);
  pindex = (int64_t *) calloc((size_t) ilen,sizeof(int64_t));
  cur = (int64_t *) calloc((size_t) nparts * (size_t) (nruns > 0 ? nruns : 1) * 2,sizeof(int64_t));
  if (pindex == NULL || beg == NULL || job == NULL || th == NULL || cur == NULL)
    { free(pindex); free(beg); free(job); free(th); free(cur); return 1; }
  fk_table_split_runs(runs,ns,nruns,tw,nparts,beg);
  for (t = 0; t < nparts; t++)
    { Ktab_Job *J = job+t;
      J->dir = dir; J->root = root; J->runs = runs; J->nruns = nruns; J->pindex = pindex;
      J->kmer = kmer; J->tw = tw; J->ib = ib; J->t = t; J->bad = 0;
      J->i = cur + (size_t) t * nruns * 2; J->j = J->i + nruns;
      for (r = 0; r < nruns; r++)
        { J->i[r] = first_byte_lower_bound(runs[r],ns[r],tw,beg[t]);
          J->j[r] = first_byte_lower_bound(runs[r],ns[r],tw,beg[t+1]);
        }
      if (pthread_create(th+t,NULL,ktab_part_thread,J) != 0) { ktab_part_thread(J); th[t] = 0; J->t = -1; }
    }
  for (t = 0; t < nparts; t++)
    { if (job[t].t >= 0) pthread_join(th[t],NULL);
      bad |= job[t].bad;
    }
  for (x = 1; x < ilen; x++) pindex[x] += pindex[x-1];
  snprintf(name,sizeof(name),"%s/%s.ktab",dir,root);
  f = open(name,O_WRONLY|O_CREAT|O_TRUNC,0700);
  if (f < 0) bad = 1;
  else
    { bad |= put(f,&kmer,sizeof(int));
      bad |= put(f,&nparts,sizeof(int));
      bad |= put(f,&cutoff,sizeof(int));
      bad |= put(f,&ib,sizeof(int));
      bad |= put(f,pindex,sizeof(int64_t)*ilen);
      close(f);
    }
  free(pindex); free(beg); free(job); free(th); free(cur);
  return bad;
}

int fk_write_ktab(const char *dir, const char *root, int kmer, int cutoff, int nparts, const uint8_t *entries, int64_t n)
{ return fk_write_ktab_runs(dir,root,kmer,cutoff,nparts,&entries,&n,1); }

/* reads back a k-mer table: the stub <dir>/<root>.ktab (kmer, nparts, cutoff, prefix bytes, cumulative prefix index) and its
   hidden parts <dir>/.<root>.ktab.<p> (kmer, n, then n x [suffix][u16 count]); the prefix of entry e is the first x with
   e < idx[x] (README.md:936-1010, the reader of libfastk.c:843-965 restated for a sequential pass)                       */
int fk_read_ktab(const char *name, int *kmer, int *cutoff, uint8_t **records, int64_t *nrec)
{ char *path = (char *) malloc(strlen(name) + 64), *dir, *root, *slash;
  int64_t *idx = NULL, ilen, e = 0, x = 0, n = 0;
  uint8_t *rec = NULL, *buf = NULL;
  int k, nparts, minval, ib, kb, tw, pw, p, f, bad = 1;
  size_t ln;
  if (path == NULL) return 1;
  strcpy(path,name);
  ln = strlen(path);
  if (ln > 5 && strcmp(path + ln - 5,".ktab") == 0) path[ln-5] = '\0';
  slash = strrchr(path,'/');
  if (slash == NULL) { dir = strdup("."); root = strdup(path); }
  else { *slash = '\0'; dir = strdup(path[0] ? path : "/"); root = strdup(slash+1); }
  { char *stub = (char *) malloc(strlen(dir) + strlen(root) + 64);
    sprintf(stub,"%s/%s.ktab",dir,root);
    f = open(stub,O_RDONLY);
    free(stub);
  }
  if (f < 0) goto done;
  if (read(f,&k,sizeof(int)) != (ssize_t) sizeof(int) || read(f,&nparts,sizeof(int)) != (ssize_t) sizeof(int)
      || read(f,&minval,sizeof(int)) != (ssize_t) sizeof(int) || read(f,&ib,sizeof(int)) != (ssize_t) sizeof(int)
      || ib < 1 || ib > 3 || k < 1 || nparts < 1)
    { close(f); goto done; }
  ilen = 1ll << (8*ib);
  idx = (int64_t *) malloc(sizeof(int64_t) * (size_t) ilen);
  if (idx == NULL || read(f,idx,sizeof(int64_t) * (size_t) ilen) != (ssize_t) (sizeof(int64_t) * (size_t) ilen)) { close(f); goto done; }
  close(f);
  n = idx[ilen-1];
  kb = (2*k+7) >> 3; tw = kb+2; pw = tw - ib;
  rec = (uint8_t *) malloc((size_t) (n > 0 ? n : 1) * tw);
  buf = (uint8_t *) malloc((size_t) pw << 16);
  if (rec == NULL || buf == NULL) goto done;
  for (p = 1; p <= nparts; p++)
    { char *part = (char *) malloc(strlen(dir) + strlen(root) + 64);
      int pk; int64_t pn, got;
      sprintf(part,"%s/.%s.ktab.%d",dir,root,p);
      f = open(part,O_RDONLY);
      free(part);
      if (f < 0) goto done;
      if (read(f,&pk,sizeof(int)) != (ssize_t) sizeof(int) || read(f,&pn,sizeof(int64_t)) != (ssize_t) sizeof(int64_t) || pk != k || e + pn > n)
        { close(f); goto done; }
      for (got = 0; got < pn; )
        { int64_t want = pn - got, i;
          ssize_t r;
          if (want > (1 << 16)) want = 1 << 16;
          r = read(f,buf,(size_t) want * pw);
          if (r != (ssize_t) (want * pw)) { close(f); goto done; }
          for (i = 0; i < want; i++, e++)
            { uint8_t *o = rec + e*tw;
              int b;
              while (x < ilen && e >= idx[x]) x++;
              if (x >= ilen) { close(f); goto done; }
              for (b = 0; b < ib; b++) o[b] = (uint8_t) (x >> (8*(ib-1-b)));
              memcpy(o+ib,buf + i*pw,(size_t) pw);
            }
          got += want;
        }
      close(f);
    }
  if (e != n) goto done;
  *kmer = k; *cutoff = minval; *records = rec; *nrec = n;
  rec = NULL;
  bad = 0;
done:
  free(path); free(dir); free(root); free(idx); free(rec); free(buf);
  return bad;
}

int64_t fk_encode_profile(const uint16_t *prof, int64_t plen, uint8_t *out)
{ uint8_t *o = out;
  int64_t  i;
  int      lz = 0;
  if (plen <= 0) return 0;
  if (prof[0] < 128) *o++ = (uint8_t) prof[0];
  else { *o++ = (uint8_t) ((prof[0] >> 8) | 0x80); *o++ = (uint8_t) (prof[0] & 0xff); }
  for (i = 1; i < plen; i++)
    { int d = (int) prof[i] - (int) prof[i-1];
      if (d == 0)
        { if (++lz >= 63) { *o++ = 63; lz = 0; }
          continue;
        }
      if (lz) { *o++ = (uint8_t) lz; lz = 0; }
      if (d >= -31 && d <= 31) *o++ = (uint8_t) (0x40 | (d & 0x3f));     /* one byte for |d| < 32, as count.c:912-913 */
      else
        { uint16_t u = (uint16_t) d;
          *o++ = (uint8_t) ((u >> 8) | 0x80);
          *o++ = (uint8_t) (u & 0xff);
        }
    }
  if (lz) *o++ = (uint8_t) lz;
  return (int64_t) (o-out);
}

typedef struct
  { const char *dir, *root; const int64_t *off; const uint16_t *prof;
    int64_t first, n; int kmer, t, bad;
  } Prof_Job;

static void *prof_part_thread(void *arg)
{ Prof_Job *J = (Prof_Job *) arg;
  const int64_t first = J->first, n = J->n;
  const int64_t *off = J->off;
  char     name[4096];
  int64_t  r, len = 0, maxp = 1;
  int64_t *idx = (int64_t *) malloc(sizeof(int64_t)*(size_t) (n > 0 ? n : 1));
  uint8_t *code;
  size_t   cap = 1 << 22, fill = 0;
  uint8_t *buf = (uint8_t *) malloc(cap);
  int      f, g;
  for (r = first; r < first+n; r++)
    if (off[r+1]-off[r] > maxp) maxp = off[r+1]-off[r];
  code = (uint8_t *) malloc((size_t) (2*maxp+4));
  snprintf(name,sizeof(name),"%s/.%s.prof.%d",J->dir,J->root,J->t+1);
  g = open(name,O_WRONLY|O_CREAT|O_TRUNC,0755);
  if (g < 0 || idx == NULL || code == NULL || buf == NULL)
    { J->bad = 1; free(idx); free(code); free(buf); if (g >= 0) close(g); return NULL; }
  for (r = first; r < first+n; r++)
    { int64_t nb = fk_encode_profile(J->prof+off[r],off[r+1]-off[r],code);
      if (fill + (size_t) nb > cap) { J->bad |= put(g,buf,fill); fill = 0; }
      if ((size_t) nb > cap) J->bad |= put(g,code,nb);
      else { memcpy(buf+fill,code,(size_t) nb); fill += nb; }
      len += nb;
      idx[r-first] = len;
    }
  J->bad |= put(g,buf,fill);
  close(g);
  snprintf(name,sizeof(name),"%s/.%s.pidx.%d",J->dir,J->root,J->t+1);
  f = open(name,O_WRONLY|O_CREAT|O_TRUNC,0755);
  if (f < 0) J->bad = 1;
  else
    { J->bad |= put(f,&J->kmer,sizeof(int));
      J->bad |= put(f,&first,sizeof(int64_t));
      J->bad |= put(f,&n,sizeof(int64_t));
      J->bad |= put(f,idx,sizeof(int64_t)*n);
      close(f);
    }
  free(idx); free(code); free(buf);
  return NULL;
}

int fk_write_prof(const char *dir, const char *root, int kmer, int nparts, const int64_t *rbeg,
                  const int64_t *off, const uint16_t *prof)
{ char name[4096];
  int  f, t, bad = 0;
  Prof_Job  *job = (Prof_Job *) calloc((size_t) (nparts > 0 ? nparts : 1),sizeof(Prof_Job));
  pthread_t *th = (pthread_t *) malloc(sizeof(pthread_t)*(size_t) (nparts > 0 ? nparts : 1));
  if (job == NULL || th == NULL) { free(job); free(th); return 1; }
  snprintf(name,sizeof(name),"%s/%s.prof",dir,root);
  f = open(name,O_WRONLY|O_CREAT|O_TRUNC,0755);
  if (f < 0) { free(job); free(th); return 1; }
  bad |= put(f,&kmer,sizeof(int));
  bad |= put(f,&nparts,sizeof(int));
  close(f);
  for (t = 0; t < nparts; t++)               /* the parts are separate files: one encoder thread each */
    { Prof_Job *J = job+t;
      J->dir = dir; J->root = root; J->off = off; J->prof = prof; J->kmer = kmer; J->t = t; J->bad = 0;
      J->first = rbeg[t]; J->n = rbeg[t+1]-rbeg[t];
      if (pthread_create(th+t,NULL,prof_part_thread,J) != 0) { prof_part_thread(J); J->t = -1; }
    }
  for (t = 0; t < nparts; t++)
    { if (job[t].t >= 0) pthread_join(th[t],NULL);
      bad |= job[t].bad;
    }
  free(job); free(th);
  return bad;
}

void fk_remove_outputs(const char *dir, const char *root)
{ /* <root>.hist / .ktab / .prof and the hidden parts .<root>.ktab.<n> / .pidx.<n> / .prof.<n>; no shell involved */
  static const char *plain[] = { "hist", "ktab", "prof", NULL };
  static const char *hidden[] = { "ktab", "pidx", "prof", NULL };
  char   name[4200], pre[4200];
  DIR   *d;
  struct dirent *e;
  int    i;
  for (i = 0; plain[i]; i++)
    { snprintf(name,sizeof(name),"%s/%s.%s",dir,root,plain[i]);
      unlink(name);
    }
  d = opendir(dir);
  if (d == NULL) return;
  while ((e = readdir(d)) != NULL)
    for (i = 0; hidden[i]; i++)
      { size_t L;
        snprintf(pre,sizeof(pre),".%s.%s.",root,hidden[i]);
        L = strlen(pre);
        if (strncmp(e->d_name,pre,L) == 0 && e->d_name[L] != '\0' && strspn(e->d_name+L,"0123456789") == strlen(e->d_name+L))
          { snprintf(name,sizeof(name),"%s/%s",dir,e->d_name);
            unlink(name);
          }
      }
  closedir(d);
}
