/*
 * mkcld -- drop-in for the reference's child-table tool (src/Mkvtree/mkcld.mn.c): `mkcld indexname` reads the lcp table
 * of an index (indexname.prj for the sizes, indexname.lcp, indexname.llv) and writes indexname.cld (Childinfo records
 * {up, down, nextlIndex}, src/include/virtualdef.h:165-179) and indexname.cld1 (one byte per entry, compresscldtab,
 * src/kurtz/mkcld.c:225-284).  The tables are computed on the GPU through the C ABI (esa_childtab); there is no CPU
 * path: without a usable device the program fails.
 *
 * Same command line, same usage text, same files, same exit codes as the reference.  Differences a user can see:
 *   - the "# overall space peak" line reports this program's host memory (the reference reports its bookkept heap and
 *     mappings, among them the suffix array, which is not needed here);
 *   - the reference maps indexname.suf (mapvirtualtreeifyoucan(SUFTAB | LCPTAB), mkcld.mn.c:58) and fails without it
 *     although it never reads it; this program asks for the file to exist, for the same exit code, and leaves it alone;
 *   - the last byte(s) of .cld1 that compresscldtab never writes hold 0 (the reference: whatever malloc returned).
 */
#include <errno.h>
#include <inttypes.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "vstree_b200.h"

static const char *progname = "mkcld";

static void *readwholefile(const char *indexname, const char *suffix, uint64_t expectbytes)
{
  char path[4096];
  FILE *fp;
  void *buf;
  size_t got;

  snprintf(path, sizeof path, "%s.%s", indexname, suffix);
  fp = fopen(path, "rb");
  if (fp == NULL)
  {
    fprintf(stderr, "%s: cannot read file \"%s\": open failed: %s\n", progname, path, strerror(errno));
    return NULL;
  }
  buf = malloc(expectbytes > 0 ? (size_t) expectbytes + 1 : 1);
  if (buf == NULL)
  {
    fprintf(stderr, "%s: not enough memory for %" PRIu64 " bytes of %s\n", progname, expectbytes, path);
    fclose(fp);
    return NULL;
  }
  got = fread(buf, 1, (size_t) expectbytes + 1, fp);  /* one more: a longer file is as wrong as a shorter one */
  fclose(fp);
  if ((uint64_t) got != expectbytes)
  {
    struct stat st;
    uint64_t have = stat(path, &st) == 0 ? (uint64_t) st.st_size : (uint64_t) got;
    /* wording of the reference: src/kurtz-basic/readvirt.c (size check of a mapped table) */
    fprintf(stderr, "%s: %" PRIu64 " bytes read from %s, but %" PRIu64 " expected\n", progname, have, path, expectbytes);
    free(buf);
    return NULL;
  }
  return buf;
}

static int writewholefile(const char *indexname, const char *suffix, const void *buf, uint64_t bytes)
{
  char path[4096];
  FILE *fp;

  snprintf(path, sizeof path, "%s.%s", indexname, suffix);
  fp = fopen(path, "wb");
  if (fp == NULL)
  {
    fprintf(stderr, "%s: cannot open file \"%s\": %s\n", progname, path, strerror(errno));
    return -1;
  }
  if (fwrite(buf, 1, (size_t) bytes, fp) != (size_t) bytes || fclose(fp) != 0)
  {
    fprintf(stderr, "%s: cannot write %" PRIu64 " bytes to \"%s\": %s\n", progname, bytes, path, strerror(errno));
    return -1;
  }
  return 0;
}

/* totallength and largelcpvalues of the project file (written by mkvtree: src/Mkvtree/mkvfile.c, read back by
   src/kurtz-basic/readvirt.c) */
static int readproject(const char *indexname, uint64_t *totallength, uint64_t *largelcpvalues)
{
  char path[4096], line[8192];
  FILE *fp;
  int have = 0;

  snprintf(path, sizeof path, "%s.prj", indexname);
  fp = fopen(path, "r");
  if (fp == NULL)
  {
    fprintf(stderr, "%s: cannot open file \"%s\": %s\n", progname, path, strerror(errno));
    return -1;
  }
  while (fgets(line, sizeof line, fp) != NULL)
  {
    if (sscanf(line, "totallength=%" SCNu64, totallength) == 1)
    {
      have |= 1;
    } else if (sscanf(line, "largelcpvalues=%" SCNu64, largelcpvalues) == 1)
    {
      have |= 2;
    }
  }
  fclose(fp);
  if (have != 3)
  {
    fprintf(stderr, "%s: file \"%s\" does not name %s\n", progname, path,
            (have & 1) ? "largelcpvalues" : "totallength");
    return -1;
  }
  return 0;
}

int main(int argc, char **argv)
{
  const char *indexname;
  uint64_t totallength = 0, nllv = 0, N;
  uint8_t *lcptab = NULL, *cldtab = NULL, *cldtab1 = NULL;
  esa_pairuint *llv = NULL;
  esa_ctx *ctx = NULL;
  char err[1024], path[4096];
  struct stat st;
  int rc = EXIT_FAILURE;

  progname = argv[0];
  if (argc != 2)
  {
    /* VSTREECHECKARGNUM, src/include/virtualdef.h:152-159 */
    fprintf(stderr, "Usage: %s %s\n", argv[0], "indexname");
    fprintf(stderr, "see Vmatch manual at http:%c%cwww.vmatch.de  for more information\n", '/', '/');
    return EXIT_FAILURE;
  }
  indexname = argv[1];
  if (readproject(indexname, &totallength, &nllv) != 0)
  {
    return EXIT_FAILURE;
  }
  N = totallength + 1;
  snprintf(path, sizeof path, "%s.suf", indexname);
  if (stat(path, &st) != 0)
  {
    fprintf(stderr, "%s: cannot read file \"%s\": open failed: %s\n", progname, path, strerror(errno));
    return EXIT_FAILURE;
  }
  lcptab = (uint8_t *) readwholefile(indexname, "lcp", N);
  if (lcptab == NULL)
  {
    goto done;
  }
  if (nllv > 0)
  {
    llv = (esa_pairuint *) readwholefile(indexname, "llv", 16 * nllv);
    if (llv == NULL)
    {
      goto done;
    }
  }
  cldtab = (uint8_t *) malloc((size_t) (3 * N));
  cldtab1 = (uint8_t *) malloc((size_t) N);
  if (cldtab == NULL || cldtab1 == NULL)
  {
    fprintf(stderr, "%s: not enough memory for the child table of %" PRIu64 " entries\n", progname, N);
    goto done;
  }
  if (esa_init(&ctx, 0, err, sizeof err) != 0
      || esa_childtab(ctx, lcptab, llv, nllv, totallength, cldtab, cldtab1, err, sizeof err) != 0)
  {
    fprintf(stderr, "%s: %s\n", progname, err);
    goto done;
  }
  if (writewholefile(indexname, "cld", cldtab, 3 * N) != 0 || writewholefile(indexname, "cld1", cldtab1, N) != 0)
  {
    goto done;
  }
  {
    const double mainbytes = 4.0 * (double) N, secondary = (double) N + 16.0 * (double) nllv;
    const double sym = totallength > 0 ? (double) totallength : 1.0;
    printf("# overall space peak: main=%.2f MB (%.2f bytes/symbol), secondary=%.2f MB (%.2f bytes/symbol)\n",
           mainbytes / (1024.0 * 1024.0), mainbytes / sym, secondary / (1024.0 * 1024.0), secondary / sym);
  }
  rc = EXIT_SUCCESS;
done:
  if (ctx != NULL)
  {
    esa_shutdown(ctx);
  }
  free(lcptab);
  free(llv);
  free(cldtab);
  free(cldtab1);
  return rc;
}
