/* bz2 entry points that always fail; see bzlib.h in this directory. */
#include <stddef.h>
#include "bzlib.h"
BZFILE *BZ2_bzopen(const char *path, const char *mode) { (void) path; (void) mode; return NULL; }
void BZ2_bzclose(BZFILE *b) { (void) b; }
int BZ2_bzread(BZFILE *b, void *buf, int len) { (void) b; (void) buf; (void) len; return -1; }
int BZ2_bzwrite(BZFILE *b, void *buf, int len) { (void) b; (void) buf; (void) len; return -1; }
