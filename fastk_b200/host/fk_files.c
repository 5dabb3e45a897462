/*  fk_files.c -- see fk_files.h.  Plain C, no CUDA. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
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

void fk_table_split(const uint8_t *entries, int64_t n, int tw, int nparts, int *beg)
{ int64_t part[256], asize = 0, sum = 0, thr, i;
  int     x, m = 0;
  memset(part,0,sizeof(part));
  for (i = 0; i < n; i++)
    part[entries[i*tw]] += tw;
  asize = n * tw;
  thr = asize / nparts;
  beg[0] = 0;
  for (x = 0; x < 256; x++)
    { sum += part[x];
      if (sum >= thr && m < nparts)
        { beg[++m] = x+1;
          thr = (asize * (m+1)) / nparts;
        }
    }
  while (m < nparts) beg[++m] = 256;
  beg[nparts] = 256;
}

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

int fk_write_ktab(const char *dir, const char *root, int kmer, int cutoff, int nparts, const uint8_t *entries, int64_t n)
{ const int kb = (2*kmer+7) >> 3, tw = kb+2;
  const int ib = fk_idx_bytes(n,kmer), pw = tw-ib;
  const int64_t ilen = 1ll << (8*ib);
  int64_t *pindex = (int64_t *) calloc((size_t) ilen,sizeof(int64_t));
  int     *beg = (int *) malloc(sizeof(int)*(nparts+1));
  char     name[4096];
  int64_t  i = 0, x;
  int      f, t, bad = 0;

  if (pindex == NULL || beg == NULL) { free(pindex); free(beg); return 1; }
  fk_table_split(entries,n,tw,nparts,beg);
  for (t = 0; t < nparts && !bad; t++)
    { int64_t j = i, m;
      size_t  cap = 1 << 22, fill = 0;
      uint8_t *buf = (uint8_t *) malloc(cap + 64);
      while (j < n && entries[j*tw] < beg[t+1]) j++;
      m = j-i;
      snprintf(name,sizeof(name),"%s/.%s.ktab.%d",dir,root,t+1);
      f = open(name,O_WRONLY|O_CREAT|O_TRUNC,0700);
      if (f < 0 || buf == NULL) { bad = 1; free(buf); break; }
      bad |= put(f,&kmer,sizeof(int));
      bad |= put(f,&m,sizeof(int64_t));
      for ( ; i < j; i++)
        { const uint8_t *e = entries + i*tw;
          int64_t idx = 0;
          int b;
          for (b = 0; b < ib; b++) idx = (idx << 8) | e[b];
          pindex[idx] += 1;
          memcpy(buf+fill,e+ib,pw);
          fill += pw;
          if (fill + pw > cap) { bad |= put(f,buf,fill); fill = 0; }
        }
      bad |= put(f,buf,fill);
      free(buf);
      close(f);
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
  free(pindex); free(beg);
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
      if (d > -32 && d < 32) *o++ = (uint8_t) (0x40 | (d & 0x3f));
      else
        { uint16_t u = (uint16_t) d;
          *o++ = (uint8_t) ((u >> 8) | 0x80);
          *o++ = (uint8_t) (u & 0xff);
        }
    }
  if (lz) *o++ = (uint8_t) lz;
  return (int64_t) (o-out);
}

int fk_write_prof(const char *dir, const char *root, int kmer, int nparts, const int64_t *rbeg,
                  const int64_t *off, const uint16_t *prof)
{ char name[4096];
  int  f, g, t, bad = 0;
  snprintf(name,sizeof(name),"%s/%s.prof",dir,root);
  f = open(name,O_WRONLY|O_CREAT|O_TRUNC,0755);
  if (f < 0) return 1;
  bad |= put(f,&kmer,sizeof(int));
  bad |= put(f,&nparts,sizeof(int));
  close(f);
  for (t = 0; t < nparts && !bad; t++)
    { int64_t r, first = rbeg[t], n = rbeg[t+1]-rbeg[t], len = 0, maxp = 1;
      int64_t *idx = (int64_t *) malloc(sizeof(int64_t)*(size_t) (n > 0 ? n : 1));
      uint8_t *code;
      size_t   cap = 1 << 22, fill = 0;
      uint8_t *buf = (uint8_t *) malloc(cap);
      for (r = first; r < first+n; r++)
        if (off[r+1]-off[r] > maxp) maxp = off[r+1]-off[r];
      code = (uint8_t *) malloc((size_t) (2*maxp+4));
      snprintf(name,sizeof(name),"%s/.%s.prof.%d",dir,root,t+1);
      g = open(name,O_WRONLY|O_CREAT|O_TRUNC,0755);
      if (g < 0 || idx == NULL || code == NULL || buf == NULL) { bad = 1; free(idx); free(code); free(buf); break; }
      for (r = first; r < first+n; r++)
        { int64_t nb = fk_encode_profile(prof+off[r],off[r+1]-off[r],code);
          if (fill + (size_t) nb > cap) { bad |= put(g,buf,fill); fill = 0; }
          if ((size_t) nb > cap) bad |= put(g,code,nb);
          else { memcpy(buf+fill,code,(size_t) nb); fill += nb; }
          len += nb;
          idx[r-first] = len;
        }
      bad |= put(g,buf,fill);
      close(g);
      snprintf(name,sizeof(name),"%s/.%s.pidx.%d",dir,root,t+1);
      f = open(name,O_WRONLY|O_CREAT|O_TRUNC,0755);
      if (f < 0) bad = 1;
      else
        { bad |= put(f,&kmer,sizeof(int));
          bad |= put(f,&first,sizeof(int64_t));
          bad |= put(f,&n,sizeof(int64_t));
          bad |= put(f,idx,sizeof(int64_t)*n);
          close(f);
        }
      free(idx); free(code); free(buf);
    }
  return bad;
}

void fk_remove_outputs(const char *dir, const char *root)
{ char cmd[8300];
  snprintf(cmd,sizeof(cmd),"rm -f %s/%s.hist %s/%s.ktab %s/.%s.ktab.* %s/%s.prof %s/.%s.pidx.* %s/.%s.prof.*",
           dir,root,dir,root,dir,root,dir,root,dir,root,dir,root);
  if (system(cmd)) {}
}
