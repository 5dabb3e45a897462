/* TEST INFRASTRUCTURE ONLY.
 *
 * Link-time stand-ins for the eleven htslib / libdeflate entry points that the reference's
 * io.c references (BAM / CRAM input only).  The oracle build (oracle/Makefile -> oracle/_ref/)
 * compiles the reference's own FastK sources where they lie under /root/reference but does NOT
 * run the vendored HTSLIB / LIBDEFLATE build systems; FASTA / FASTQ input (all that the parity
 * suite feeds it) never reaches these symbols.  Anything that does reach one aborts loudly.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>

static void nope(const char *what)
{ fprintf(stderr, "oracle/_ref: %s is not available in the oracle build (BAM/CRAM unsupported)\n", what);
  abort();
}

void *cram_open(const char *path, const char *mode) { (void) path; (void) mode; nope("cram_open"); return NULL; }
int   cram_close(void *fd) { (void) fd; nope("cram_close"); return -1; }
void *cram_get_seq(void *fd) { (void) fd; nope("cram_get_seq"); return NULL; }
int   hgetc2(void *fp) { (void) fp; nope("hgetc2"); return -1; }
long  hread2(void *fp, void *buf, size_t n, size_t m) { (void) fp; (void) buf; (void) n; (void) m; nope("hread2"); return -1; }
long  hseek(void *fp, long off, int whence) { (void) fp; (void) off; (void) whence; nope("hseek"); return -1; }
int   itf8_decode(void *fd, int *val) { (void) fd; (void) val; nope("itf8_decode"); return -1; }

/* io.c allocates one decompressor per input thread up front, whatever the file type, so these
 * two must succeed; the decompress calls themselves are only reached for BAM input.           */
void *libdeflate_alloc_decompressor(void) { return malloc(16); }
void  libdeflate_free_decompressor(void *d) { free(d); }
unsigned libdeflate_crc32(unsigned crc, const void *buf, size_t len) { (void) crc; (void) buf; (void) len; nope("libdeflate_crc32"); return 0; }
int   libdeflate_gzip_decompress(void *d, const void *in, size_t inb, void *out, size_t outb, size_t *act)
{ (void) d; (void) in; (void) inb; (void) out; (void) outb; (void) act; nope("libdeflate_gzip_decompress"); return -1; }
