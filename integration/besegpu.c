/*
 * besegpu.c -- drop-in replacement of the reference's suffix sorter AT ITS OWN SEAM.
 *
 * The reference sorts in src/Mkvtree/bese.c (compiled twice: besesuffixsort, and with -DSPECIAL besesuffixsortspecial)
 * on top of the bucket sort of src/Mkvtree/ppsort.c; mkvtreeprocess (src/Mkvtree/mkvprocess.c:976-1006) calls one of the
 * two and then derives every other table from what they leave behind in MKVaux / Virtualtree.  This file defines the
 * same two functions, with the same signature and the same post-conditions, on top of libvstree_b200 (esa_build).
 * Linked INSTEAD of bese.o / besespecial.o / ppsort.o, with every other object of the reference unmodified, it turns
 *   mkvtree, mkvram, mkrcidx, mkdna6idx   (callmkvtree / callmkvtreegeneric, src/Mkvtree/mkvtree.c:589-707)
 *   vmatch's on-the-fly query index        (completevirtualtree, mkvtree.c:709-738, src/Vmatch/procmatch.c:68-79)
 * into GPU builds: SURVEY.md section 8f row 3.  oracle/build_ref.sh builds those binaries next to the reference's own
 * (oracle/_ref/gpu/, reference objects + this file); tests/test_gpu_refseam.py compares the two families byte for byte.
 *
 * Compiled against the REFERENCE's headers (where /root/reference is present); nothing of the product depends on it.
 *
 * Post-conditions reproduced (what mkvprocess.c consumes afterwards):
 *   mkvaux->sortedsuffixes     Uint[totallength+1], the suffix array (becomes virtualtree->suftab, mkvprocess.c:1017-1023)
 *   lcp values                 totallength+1 bytes into lcpfp (storeonfile) or virtualtree->lcptab (outlcpsubtab, bese.c:533-591)
 *   virtualtree->largelcpvalues, ->maxbranchdepth                          (bese.c:553-566)
 *   virtualtree->numofcodes, mkvaux->leftborder[numofcodes+1], the per-bucket special counts as a sorted code/distance
 *   list (mkvaux->arcodedistance, storecodes = True)                       (ppsort.c:93-230, read by makebcktab :251-327)
 *   mkvaux->lcpsubtab          freed by freeeverything (mkvprocess.c:818-832)
 *   the three -v lines of bese.c:871,901,915 / :1051
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "types.h"
#include "debugdef.h"
#include "spacedef.h"
#include "arraydef.h"
#include "errordef.h"
#include "virtualdef.h"
#include "fhandledef.h"
#include "rmnsufinfo.h"
#include "mkvaux.h"

#include "detnumofcodes.pr"

#include "vstree_b200.h"

static esa_ctx *esactx = NULL;   /* one context per process, like the reference's other static registries */

static Sint gpusort(MKVaux *mkvaux, Virtualtree *virtualtree, Uint numofchars, BOOL demandlcp, FILE *lcpfp,
                    Showverbose showverbose, const DefinedUint *maxdepth)
{
  const Uint totallength = virtualtree->multiseq.totallength, N = totallength + 1;
  const Uint k = virtualtree->prefixlength;
  esa_result res;
  char err[512];
  uint32_t demand = ESA_SUFTAB;
  Uint i;

  if(maxdepth->defined)
  {
    ERROR0("option -maxdepth (remainsort.c) is not available with the GPU suffix sorter");
    return (Sint) -1;
  }
  if(showverbose != NULL)
  {
    showverbose("initializing data structures");
  }
  if(esactx == NULL && esa_init(&esactx, 0, err, sizeof err) != 0)
  {
    ERROR1("%s", err);           /* no GPU: fail loudly, there is no CPU path behind this seam */
    return (Sint) -2;
  }
  if(demandlcp)
  {
    demand |= ESA_LCPTAB;
  }
  if(k > 0)
  {
    char sbuf[256];
    demand |= ESA_BCKTAB;        /* leftborder and the special counts come out of the bucket table */
    if(showverbose != NULL)
    {
      sprintf(sbuf, "sorting suffixes according to prefix of length %lu", (Showuint) k);
      showverbose(sbuf);
    }
  }
  if(showverbose != NULL)
  {
    showverbose(k > 0 ? "sorting all buckets" : "sorting suffixes");
  }
  if(esa_build(esactx, virtualtree->multiseq.sequence, (uint64_t) totallength, (uint32_t) numofchars, (uint32_t) k,
               demand, &res, NULL, err, sizeof err) != 0)
  {
    ERROR1("%s", err);
    return (Sint) -3;
  }
  /* the suffix array, in the reference's own bookkept memory (Uint is 8 bytes little endian: types.h:41-49) */
  ALLOCASSIGNSPACE(mkvaux->sortedsuffixes, NULL, Suffixptr, N);
  memcpy(mkvaux->sortedsuffixes, res.suftab, sizeof(Uint) * (size_t) N);
  mkvaux->sentinel = (Suffixptr) totallength;
  ALLOCASSIGNSPACE(mkvaux->lcpsubtab, NULL, Uint, UintConst(1));
  /* lcp bytes, large values, maximal branching depth */
  INITARRAY(&virtualtree->largelcpvalues, PairUint);
  virtualtree->maxbranchdepth = 0;
  if(demandlcp)
  {
    if(lcpfp != NULL)
    {
      if(fwrite(res.lcptab, sizeof(Uchar), (size_t) N, lcpfp) != (size_t) N)
      {
        ERROR0("cannot write the lcp table");
        esa_free(esactx, &res);
        return (Sint) -4;
      }
    } else
    {
      memcpy(virtualtree->lcptab, res.lcptab, (size_t) N);
    }
    virtualtree->maxbranchdepth = (Uint) res.maxbranchdepth;
    if(res.nllv > 0)
    {
      ALLOCASSIGNSPACE(virtualtree->largelcpvalues.spacePairUint, NULL, PairUint, (Uint) res.nllv);
      memcpy(virtualtree->largelcpvalues.spacePairUint, res.llv, sizeof(PairUint) * (size_t) res.nllv);
      virtualtree->largelcpvalues.allocatedPairUint = virtualtree->largelcpvalues.nextfreePairUint = (Uint) res.nllv;
    }
  }
  /* what makebcktab reads: bucket starts and, per bucket, the number of suffixes with a special inside the prefix */
  if(k > 0)
  {
    Uint specials = 0, j = 0;
    virtualtree->numofcodes = vm_determinenumofcodes(numofchars, k);
    ALLOCASSIGNSPACE(mkvaux->leftborder, NULL, Uint, virtualtree->numofcodes + 1);
    for(i = 0; i < virtualtree->numofcodes; i++)
    {
      const Uint right = (i + 1 < virtualtree->numofcodes) ? (Uint) res.bcktab[2 * (i + 1)] : N;
      mkvaux->leftborder[i] = (Uint) res.bcktab[2 * i];
      specials += right - (Uint) res.bcktab[2 * i + 1];
    }
    mkvaux->leftborder[virtualtree->numofcodes] = N;
    mkvaux->storecodes = True;
    mkvaux->specialtable = NULL;
    INITARRAY(&mkvaux->arcodedistance, Uint);
    ALLOCASSIGNSPACE(mkvaux->arcodedistance.spaceUint, NULL, Uint, specials + 1);
    for(i = 0; i < virtualtree->numofcodes; i++)
    {
      const Uint right = (i + 1 < virtualtree->numofcodes) ? (Uint) res.bcktab[2 * (i + 1)] : N;
      Uint c = right - (Uint) res.bcktab[2 * i + 1];
      for(/* Nothing */; c > 0; c--)
      {
        mkvaux->arcodedistance.spaceUint[j++] = CODEDISTANCE(i, 0);
      }
    }
    mkvaux->arcodedistance.spaceUint[j] = CODEDISTANCE(virtualtree->numofcodes, 0);  /* ends the walk of makebcktab */
    mkvaux->arcodedistance.allocatedUint = specials + 1;
    mkvaux->arcodedistance.nextfreeUint = specials;
    mkvaux->maxbucketsize = 0;
  }
  esa_free(esactx, &res);
  return 0;
}

Sint besesuffixsort(MKVaux *mkvaux, Virtualtree *virtualtree, Uint numofchars, BOOL demandlcp, FILE *lcpfp,
                    Showverbose showverbose, const DefinedUint *maxdepth)
{
  return gpusort(mkvaux, virtualtree, numofchars, demandlcp, lcpfp, showverbose, maxdepth);
}

Sint besesuffixsortspecial(MKVaux *mkvaux, Virtualtree *virtualtree, Uint numofchars, BOOL demandlcp, FILE *lcpfp,
                           Showverbose showverbose, const DefinedUint *maxdepth)
{
  return gpusort(mkvaux, virtualtree, numofchars, demandlcp, lcpfp, showverbose, maxdepth);
}
