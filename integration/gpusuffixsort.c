/*
 * gpusuffixsort.c -- the reference-side binding of INTEGRATION.md section B as a compilable unit.
 *
 * This is what a maintainer of the reference adds next to src/Mkvtree/mkvprocess.c to route the construction block of
 * mkvtreeprocess (mkvprocess.c:949-1060) through libvstree_b200.  It is compiled against the REFERENCE's own headers
 * (types.h, virtualdef.h, inputdef.h, errordef.h ...) by tests/test_integration_binding.py when the reference tree is
 * mounted; nothing of the product depends on it.
 */
#include "types.h"
#include "errordef.h"
#include "virtualdef.h"
#include "inputdef.h"

#include "vstree_b200.h"          /* include/vstree_b200.h of this repository */

static esa_ctx *esactx = NULL;    /* one context per process, like the reference's other static registries */

static Sint gpusuffixsort(Input *input, Virtualtree *virtualtree)
{
  esa_result res;
  char err[512];
  uint32_t demand = 0;

  if(esactx == NULL && esa_init(&esactx, 0, err, sizeof err) != 0)
  {
    ERROR1("%s", err);            /* no GPU: fail loudly, there is no CPU path behind this call */
    return (Sint) -1;
  }
  /* TISTAB..STI1TAB have the same bit values as ESA_* (src/include/virtualdef.h:24-64) */
  demand = (uint32_t) (input->demand & (BWTTAB | SUFTAB | LCPTAB | BCKTAB | STI1TAB));
  demand |= ESA_SUFTAB;           /* suftab is always produced (longest, sti1, bwt need it) */
  if(esa_build(esactx,
               virtualtree->multiseq.sequence,             /* tis: codes, 254 wildcard, 255 separator */
               (uint64_t) virtualtree->multiseq.totallength,
               (uint32_t) input->numofchars,
               (uint32_t) virtualtree->prefixlength,
               demand, &res, NULL, err, sizeof err) != 0)
  {
    ERROR1("%s", err);
    return (Sint) -2;
  }
  /* Uint is 8 bytes little endian here (src/include/types.h:41-49), so the tables can be taken as they are.
     The library owns the pinned buffers until esa_free(); a maintainer either copies them into
     ALLOCASSIGNSPACE'd arrays (in-memory use, completevirtualtree) or writes them out and frees (storeonfile). */
  virtualtree->suftab = (Uint *) res.suftab;                                 /* besesuffixsort, bese.c:855 */
  virtualtree->lcptab = (Uchar *) res.lcptab;                                /* outlcpsubtab, bese.c:533 */
  virtualtree->largelcpvalues.spacePairUint = (PairUint *) res.llv;          /* PairUint == esa_pairuint */
  virtualtree->largelcpvalues.nextfreePairUint = (Uint) res.nllv;
  virtualtree->maxbranchdepth = (Uint) res.maxbranchdepth;
  virtualtree->longest.defined = True;                                       /* determinelongest, :857 */
  virtualtree->longest.uintvalue = (Uint) res.longest;
  virtualtree->bcktab = (Uint *) res.bcktab;                                 /* makebcktab, :251 */
  virtualtree->numofcodes = (Uint) res.numofcodes;
  virtualtree->stitab1 = (Uchar *) res.stitab1;                              /* makestitab1, :583 */
  virtualtree->bwttab = (Uchar *) res.bwttab;                                /* makebwttab, :354 */
  virtualtree->constructed |= (input->demand & (BWTTAB | SUFTAB | LCPTAB | BCKTAB | STI1TAB));
  return 0;
}
