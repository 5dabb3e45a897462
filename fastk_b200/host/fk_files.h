/*  fk_files.h -- host-side writers of FastK's three result formats (.hist / .ktab / .prof), fed from
 *  the fkgpu_result of libfastk_gpu.  Byte layouts follow the reference writers:
 *     .hist   count.c:1893-1910        .ktab  table.c:216-217,282-284,329-333,483-498
 *     .prof   merge.c:871-872,926-928,977-979, profile code count.c:886-921 + merge.c:541-562
 *  (README.md:936-1069 is the normative format description).
 */
#ifndef FK_FILES_H
#define FK_FILES_H
#include <stdint.h>

/* IDX_BYTES rule of count.c:1620-1626 */
int  fk_idx_bytes(int64_t nentries, int kmer);

/* first-byte cut points for nparts table parts, rule of MSDsort.c:330-352 applied to the entry bytes */
void fk_table_split(const uint8_t *entries, int64_t n, int tmer_word, int nparts, int *beg /*[nparts+1]*/);

/* the same over a table delivered as nruns sorted runs with disjoint keys (fkgpu_result.run_table / run_ntable) */
void fk_table_split_runs(const uint8_t *const *runs, const int64_t *ns, int nruns, int tmer_word, int nparts, int *beg);

int  fk_write_hist(const char *dir, const char *root, int kmer, const int64_t *hist /*[32768]*/, int64_t max_inst);

/* entries = n records [kmer_bytes key][u16 LE count], sorted */
int  fk_write_ktab(const char *dir, const char *root, int kmer, int cutoff, int nparts,
                   const uint8_t *entries, int64_t n);

/* the table as nruns sorted runs with disjoint keys: every part merges its slices of the runs (table.c:240-313) */
int  fk_write_ktab_runs(const char *dir, const char *root, int kmer, int cutoff, int nparts,
                        const uint8_t *const *runs, const int64_t *ns, int nruns);

/* a k-mer table read back into n records [kmer_bytes key][u16 LE count] in key order (malloc'ed: the caller frees);
   name = <dir>/<root> with or without the .ktab extension (as Open_Kmer_Stream takes it, libfastk.c:843).  0 = ok. */
int  fk_read_ktab(const char *name, int *kmer, int *cutoff, uint8_t **records, int64_t *n);

/* greedy profile code of one read's count vector; returns # of bytes written (out needs 2*plen+2) */
int64_t fk_encode_profile(const uint16_t *prof, int64_t plen, uint8_t *out);

/* nparts profile parts; part t holds reads [rbeg[t], rbeg[t+1]) of the global order; off/prof as returned
 * by fkgpu_profiles                                                                                    */
int  fk_write_prof(const char *dir, const char *root, int kmer, int nparts, const int64_t *rbeg,
                   const int64_t *off, const uint16_t *prof);

/* remove <dir>/<root>.{hist,ktab,prof} and their hidden parts (Clean_Exit, FastK.c:181-221) */
void fk_remove_outputs(const char *dir, const char *root);

#endif
