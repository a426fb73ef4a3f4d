/*
 * mkcldgpu.c -- the reference's child-table functions AT THEIR OWN SEAM, on top of libvstree_b200 (esa_childtab).
 *
 * The reference builds the child table in src/kurtz/mkcld.c: mkcldtabnextlIndex (:41-108), mkcldtabupdown (:110-207)
 * and compresscldtab (:225-284), called one after the other by makecldtabandoutput (src/Mkvtree/mkcld.mn.c:13-47).
 * This file defines the same three functions with the same signatures.  Linked INSTEAD of mkcld.o of libkurtz, with
 * every other object of the reference unmodified, it turns the reference's mkcld into a GPU build: oracle/build_ref.sh
 * links oracle/_ref/gpu/mkcld.x that way and tests/test_gpu_childtab.py compares its files with those of the unmodified
 * oracle/_ref/mkcld.x.  SURVEY.md section 8f row 4; INTEGRATION.md section B3.
 *
 * One device pass computes all fields, so the first call does the work and keeps the other two results until the
 * functions that deliver them are called (the reference calls all three, in this order, on the same tables).
 *
 * Compiled against the REFERENCE's headers (where /root/reference is present); nothing of the product depends on it.
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
#include "failures.h"

#include "vstree_b200.h"

static esa_ctx *esactx = NULL;
static Uchar *pending = NULL;      /* 3 * (totallength + 1) record bytes, then totallength + 1 bytes of the small table */
static Uint pendinglength = 0;

static void gpuchildtab(Uchar *lcptab, ArrayPairUint *largelcpvalues, Uint totallength)
{
  char err[512];
  const Uint N = totallength + 1;

  if(esactx == NULL && esa_init(&esactx, 0, err, sizeof err) != 0)
  {
    fprintf(stderr, "%s\n", err);   /* no GPU: fail loudly, there is no CPU path behind this seam */
    exit(EXIT_FAILURE);
  }
  free(pending);
  pending = (Uchar *) malloc((size_t) (4 * N));
  if(pending == NULL)
  {
    fprintf(stderr, "not enough memory for the child table of %lu entries\n", (Showuint) N);
    exit(EXIT_FAILURE);
  }
  /* PairUint is {Uint uint0, uint1} with 8-byte Uint (types.h:41-49): the layout of esa_pairuint */
  if(esa_childtab(esactx, lcptab, (const esa_pairuint *) largelcpvalues->spacePairUint,
                  (uint64_t) largelcpvalues->nextfreePairUint, (uint64_t) totallength, pending, pending + 3 * N,
                  err, sizeof err) != 0)
  {
    fprintf(stderr, "%s\n", err);
    exit(EXIT_FAILURE);
  }
  pendinglength = totallength;
}

void mkcldtabnextlIndex(Childinfo *cldtab, Uchar *lcptab, ArrayPairUint *largelcpvalues, Uint totallength)
{
  Uint i;

  gpuchildtab(lcptab, largelcpvalues, totallength);
  for(i = 0; i <= totallength; i++)
  {
    cldtab[i].nextlIndex = pending[3 * i + 2];
  }
}

void mkcldtabupdown(Childinfo *cldtab, Uchar *lcptab, ArrayPairUint *largelcpvalues, Uint totallength)
{
  Uint i;

  if(pending == NULL || pendinglength != totallength)
  {
    gpuchildtab(lcptab, largelcpvalues, totallength);
  }
  for(i = 0; i <= totallength; i++)
  {
    cldtab[i].up = pending[3 * i];
    cldtab[i].down = pending[3 * i + 1];
  }
}

void compresscldtab(Virtualtree *virtualtree, Uchar *cldtab1)
{
  const Uint totallength = virtualtree->multiseq.totallength;

  if(pending == NULL || pendinglength != totallength)
  {
    gpuchildtab(virtualtree->lcptab, &virtualtree->largelcpvalues, totallength);
  }
  memcpy(cldtab1, pending + 3 * (totallength + 1), (size_t) (totallength + 1));
  free(pending);
  pending = NULL;
}
