/* Stand-in for <bzlib.h>, needed only so that the reference's genfile.c compiles
   when oracle/build_ref.sh builds the reference mkvtree as the parity checker.
   bz2 input is never exercised by mkvtree. Test infrastructure, not product code. */
#ifndef ORACLE_STUB_BZLIB_H
#define ORACLE_STUB_BZLIB_H
typedef void BZFILE;
BZFILE *BZ2_bzopen(const char *path, const char *mode);
void BZ2_bzclose(BZFILE *b);
int BZ2_bzread(BZFILE *b, void *buf, int len);
int BZ2_bzwrite(BZFILE *b, void *buf, int len);
#endif
