/* TEST INFRASTRUCTURE ONLY -- fastk_oracle: a plain-C, single-threaded CPU restatement of the
 * FastK k-mer counting hot path (encode + canonicalise -> sort -> run-length count -> histogram ->
 * table / profile emit).  It exists to CHECK the CUDA path; nothing under fastk_b200/ may call,
 * link or import it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
 *
 * Parity pin: the reference ships no golden vectors, so this restatement is pinned
 * against outputs of the reference itself, built from its own sources into oracle/_ref/ (see
 * oracle/Makefile) -- tests/test_oracle_vs_ref.py and the committed fixtures in tests/golden/.
 *
 * Every function cites the reference file:line (paths relative to /root/reference) it restates.
 */
#ifndef FASTK_ORACLE_H
#define FASTK_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
  { int      k;        /* k-mer length                                              */
    int      kbytes;   /* (2k+7)>>3, FastK.c:419                                    */
    int64_t  n;        /* # of distinct canonical k-mers                            */
    uint8_t *keys;     /* n * kbytes, strictly increasing bytewise                  */
    int64_t *cnt;      /* true (unsaturated) # of instances of each                 */
  } FKO_Table;

/* reads are given DATA_BLOCK style (FastK.h:87-98): read i is bases[boff[i] .. boff[i+1]-2], each
 * followed by one terminator byte.  bc_prefix = -bc<n> (split.c:1075).                         */
FKO_Table *fko_count(const char *bases, const int64_t *boff, int64_t nreads, int k, int bc_prefix);
void       fko_free_table(FKO_Table *t);

/* hist[0] unused, hist[1..32767]; hist[32767] counts k-mers with >= 32767 instances and *max_inst is the
 * sum of their true instance counts (MSDsort.c:491-509, count.c:455-458,1543-1553).                   */
void       fko_histogram(const FKO_Table *t, int64_t *hist /*[32768]*/, int64_t *max_inst);

/* table records [kbytes key][u16 LE min(cnt,32767)] for every k-mer with cnt >= cutoff
 * (count.c:564-616); out may be NULL to just count.  Returns # of records.                  */
int64_t    fko_table_entries(const FKO_Table *t, int cutoff, uint8_t *out);

/* count profile of one read (Appendix B of SURVEY.md; count.c:868-947): prof[i] = min(cnt,32767) of the
 * canonical k-mer at position i of the read (after the bc prefix), 0 if it covers a non-acgt char.
 * Returns the profile length max(0, len-bc-k+1).                                                     */
int64_t    fko_profile(const FKO_Table *t, const char *seq, int64_t len, int bc_prefix, uint16_t *prof);

/* greedy canonical profile code (merge.c:534-716, count.c:886-921) and its decoder (libfastk.c:1707-1803) */
int64_t    fko_encode_profile(const uint16_t *prof, int64_t plen, uint8_t *out);
int64_t    fko_decode_profile(const uint8_t *code, int64_t nbytes, uint16_t *prof, int64_t cap);

/* IDX_BYTES rule, count.c:1620-1626 */
int        fko_idx_bytes(int64_t nentries, int k);

/* thread/part split rule on first key byte, MSDsort.c:330-352: beg[0..nparts] first-byte cut points */
void       fko_part_split(const int64_t part[256], int nparts, int *beg /*[nparts+1]*/);

/* file writers: byte layouts of Appendix A (count.c:1893-1910, table.c:216-217,282-284,483-498,
 * merge.c:871-872,926-928,977-979).  Return 0 on success.                                       */
int fko_write_hist(const char *dir, const char *root, int k, const int64_t *hist, int64_t max_inst);
int fko_write_ktab(const char *dir, const char *root, int k, int cutoff, int nparts,
                   const uint8_t *entries, int64_t nentries);
int fko_write_prof(const char *dir, const char *root, int k, int nparts, const int64_t *part_reads,
                   const uint8_t *const *part_code, const int64_t *const *part_off);

/* whole FastK run over FASTA/FASTQ files -> <dir>/<root>.{hist,ktab,prof}; the input automaton
 * restates io.c:678-734.  do_table = -t cutoff (0 = none), do_profile = -p.                  */
int fko_run_files(int nfiles, char **files, const char *dir, const char *root,
                  int k, int do_table, int do_profile, int bc_prefix, int compress, int nparts);

#ifdef __cplusplus
}
#endif
#endif
