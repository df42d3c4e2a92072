/* ORACLE / TEST INFRASTRUCTURE -- not product code.
 *
 * The reference links htslib only for its `-bam` writer (reference src/ReadMapping.cpp:92-122,
 * 550-557, 701).  BAM output is outside the hot path, so instead of
 * building the vendored htslib (needs bz2/lzma headers that this image lacks) the oracle build
 * satisfies the 8 imported symbols with stubs that abort when reached.  SAM text and VCF output,
 * which the parity tests use, never touch them.
 */
#include <stdio.h>
#include <stdlib.h>

static void die(const char *fn) {
    fprintf(stderr, "[oracle] %s: BAM output is not available in the oracle build (use -sam)\n", fn);
    abort();
}
void *bam_init1(void) { die("bam_init1"); return 0; }
void bam_destroy1(void *b) { (void)b; die("bam_destroy1"); }
int hts_close(void *fp) { (void)fp; die("hts_close"); return -1; }
void *hts_open_format(const char *fn, const char *mode, const void *fmt) {
    (void)fn; (void)mode; (void)fmt; die("hts_open_format"); return 0;
}
void *sam_hdr_parse(int l_text, const char *text) { (void)l_text; (void)text; die("sam_hdr_parse"); return 0; }
int sam_hdr_write(void *fp, const void *h) { (void)fp; (void)h; die("sam_hdr_write"); return -1; }
int sam_parse1(void *s, void *h, void *b) { (void)s; (void)h; (void)b; die("sam_parse1"); return -1; }
int sam_write1(void *fp, const void *h, const void *b) { (void)fp; (void)h; (void)b; die("sam_write1"); return -1; }
