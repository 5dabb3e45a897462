/* TEST INFRASTRUCTURE ONLY -- command-line front end of the CPU oracle (see fastk_oracle.h).
 * Accepts the FastK option grammar subset -k<int> -t[<int>] -p -c -bc<int> -T<int> -N<path> (FastK.c:250-326)
 * and writes <root>.hist / .ktab / .prof next to the first input (or at -N).                          */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <libgen.h>
#include "fastk_oracle.h"

int main(int argc, char *argv[])
{ int   k = 40, table = 0, prof = 0, bc = 0, hoco = 0, T = 4;
  char *out = NULL;
  char *files[1024];
  int   nfiles = 0, i;
  char  dirbuf[4096], rootbuf[4096];

  for (i = 1; i < argc; i++)
    if (argv[i][0] == '-')
      { char *a = argv[i];
        if (a[1] == 'k') k = atoi(a+2);
        else if (a[1] == 'T') T = atoi(a+2);
        else if (a[1] == 'N') out = a+2;
        else if (a[1] == 'b' && a[2] == 'c') bc = atoi(a+3);
        else if (a[1] == 't' && isdigit((unsigned char) a[2])) table = atoi(a+2);
        else if (a[1] == 'P' || a[1] == 'M') ;
        else
          { char *c;
            for (c = a+1; *c; c++)
              if (*c == 't') table = table ? table : 1;
              else if (*c == 'p') prof = 1;
              else if (*c == 'c') hoco = 1;
              else if (*c == 'v') ;
              else { fprintf(stderr,"fastk_oracle: unknown option %s\n",a); return (1); }
          }
      }
    else if (nfiles < 1024)
      files[nfiles++] = argv[i];
  if (nfiles == 0)
    { fprintf(stderr,"Usage: fastk_oracle [-k<int>] [-t[<int>]] [-p] [-c] [-bc<int>] [-T<int>] [-N<out>] <fasta|fastq> ...\n");
      return (1);
    }
  { char *src = strdup(out ? out : files[0]);
    char *s2  = strdup(src);
    char *dot;
    strcpy(dirbuf,dirname(src));
    strcpy(rootbuf,basename(s2));
    if (!out)
      { static const char *sfx[] = { ".fasta", ".fastq", ".fa", ".fq", NULL };
        int j;
        for (j = 0; sfx[j]; j++)
          { size_t L = strlen(rootbuf), S = strlen(sfx[j]);
            if (L > S && strcmp(rootbuf+L-S,sfx[j]) == 0)
              { rootbuf[L-S] = '\0'; break; }
          }
      }
    (void) dot;
    free(src); free(s2);
  }
  return (fko_run_files(nfiles,files,dirbuf,rootbuf,k,table,prof,bc,hoco,T));
}
