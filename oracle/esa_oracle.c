/*
 * esa_oracle.c -- CPU restatement ("port" oracle) of mkvtree's enhanced-suffix-array construction path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (vstree_b200/) includes, links or executes this file;
 * it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA path.
 *
 * Parity status: PINNED.  tests/test_oracle.py compares every table this file produces, byte for byte,
 * with the output of the unmodified reference binary (oracle/_ref/mkvtree.x, built by oracle/build_ref.sh
 * from /root/reference/src) on the reference's own fixtures and on the committed goldens in tests/golden/.
 *
 * It follows the reference's structure (bucket sort on the padded k-mer code, then a multikey quicksort per
 * bucket, LCP by direct comparison, derived tables by linear walks) but is written from scratch; each
 * function names the reference lines whose behaviour it restates.  Plain C99, single threaded.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORACLE_WILDCARD 254u      /* reference: src/include/chardef.h:19-43 (WILDCARD, SEPARATOR, ISSPECIAL) */
#define ORACLE_SEPARATOR 255u
#define ORACLE_UNDEFBWT 253u      /* reference: src/include/chardef.h:50 (UNDEFBWTCHAR) */
#define IS_SPECIAL(c) ((c) >= ORACLE_WILDCARD)

typedef struct
{
  const uint8_t *tis;
  uint64_t n;
} Text;

/* Symbol order.  reference: src/Mkvtree/bese.c:26-55 (UNIQUEINT, GENDEREF, ISNOTEND).
   A letter compares by its code; a special at position p, and the sentinel at p == n, compares as p+256,
   i.e. above every letter and different from every other special. */
static inline uint64_t sym(const Text *t, uint64_t p)
{
  if (p < t->n && !IS_SPECIAL(t->tis[p]))
  {
    return t->tis[p];
  }
  return p + 256u;
}

/* reference: src/kurtz-basic/detnumofcodes.c:8-11 -- same libm expression */
uint64_t oracle_numofcodes(uint32_t numofchars, uint32_t prefixlength)
{
  return (uint64_t) pow((double) numofchars, (double) prefixlength);
}

/* reference: src/kurtz/detpfxlen.c:31-50 (logalphasize) */
static uint64_t logalphasize(uint64_t numofchars, double value)
{
  if (value <= (double) numofchars)
  {
    return 1;
  }
  return (uint64_t) floor(log(value) / log((double) numofchars));
}

/* reference: src/kurtz/detpfxlen.c:52-61; sizeofbckentry is 2*sizeof(Uint) = 16 at the call site mkvtree.c:652-658 */
uint32_t oracle_recommendedprefixlength(uint32_t numofchars, uint64_t totallength)
{
  uint64_t k = logalphasize(numofchars, (double) totallength / 16);
  return (uint32_t) (k == 0 ? 1 : k);
}

/* reference: src/kurtz/detpfxlen.c:63-89 with PREFIXLENBITS = 4 (src/Mkvtree/mkvaux.h:27) and
   MAXMULTIPLIEROFTOTALLENGTH = 4 */
uint32_t oracle_maximalprefixlength(uint32_t numofchars, uint64_t totallength)
{
  const unsigned prefixlenbits = 4;
  uint64_t maxk = logalphasize(numofchars, (double) totallength / (16 / 4));
  /* MAXREMAININGAFTERPREFIXLEN(bits) = 2^(64-bits)-1, MAXVALUEWITHBITS(bits) = 2^bits-1 (src/include/intbits.h) */
  uint64_t tmp = logalphasize(numofchars, (double) ((((uint64_t) 1) << (64 - prefixlenbits)) - 1));
  if (maxk > tmp)
  {
    maxk = tmp;
  }
  tmp = (((uint64_t) 1) << prefixlenbits) - 1;
  if (maxk > tmp)
  {
    maxk = tmp;
  }
  return (uint32_t) maxk;
}

/* Longest common prefix of two suffixes under the symbol order above.
   reference: src/Mkvtree/bese.c:91-95 (GETLCP): specials never match, so the loop stops there. */
static uint64_t suffix_lcp(const Text *t, uint64_t a, uint64_t b, uint64_t from)
{
  uint64_t l = from;
  while (sym(t, a + l) == sym(t, b + l))
  {
    l++;
  }
  return l;
}

/* ---- multikey quicksort on one bucket.  reference: src/Mkvtree/bese.c:643-853 (medianof3, inssort,
   bentleysedgewick).  Only the resulting order matters (it is a total order, so any correct sorter
   gives the same table); this one is a textbook ternary-split quicksort with an insertion-sort base. ---- */

static int suffix_cmp(const Text *t, uint64_t a, uint64_t b, uint64_t depth)
{
  for (;; depth++)
  {
    uint64_t ca = sym(t, a + depth), cb = sym(t, b + depth);
    if (ca != cb)
    {
      return ca < cb ? -1 : 1;
    }
  }
}

static void insertion_sort(const Text *t, uint64_t *a, int64_t cnt, uint64_t depth)
{
  for (int64_t i = 1; i < cnt; i++)
  {
    uint64_t x = a[i];
    int64_t j = i;
    while (j > 0 && suffix_cmp(t, a[j - 1], x, depth) > 0)
    {
      a[j] = a[j - 1];
      j--;
    }
    a[j] = x;
  }
}

static inline void swap64(uint64_t *a, int64_t i, int64_t j)
{
  uint64_t tmp = a[i];
  a[i] = a[j];
  a[j] = tmp;
}

static uint64_t median3(uint64_t a, uint64_t b, uint64_t c)
{
  if (a < b)
  {
    return b < c ? b : (a < c ? c : a);
  }
  return a < c ? a : (b < c ? c : b);
}

static void multikey_qsort(const Text *t, uint64_t *a, int64_t cnt, uint64_t depth)
{
  while (cnt > 1)
  {
    if (cnt <= 8)
    {
      insertion_sort(t, a, cnt, depth);
      return;
    }
    uint64_t pivot = median3(sym(t, a[0] + depth), sym(t, a[cnt / 2] + depth), sym(t, a[cnt - 1] + depth));
    int64_t lt = 0, gt = cnt - 1, i = 0;
    while (i <= gt)
    {
      uint64_t c = sym(t, a[i] + depth);
      if (c < pivot)
      {
        swap64(a, lt++, i++);
      } else if (c > pivot)
      {
        swap64(a, i, gt--);
      } else
      {
        i++;
      }
    }
    multikey_qsort(t, a, lt, depth);
    multikey_qsort(t, a + gt + 1, cnt - gt - 1, depth);
    /* equal part: continue one symbol deeper (a pivot >= 256 is unique, so that part then has size 1) */
    a += lt;
    cnt = gt - lt + 1;
    depth++;
  }
}

/*
 * Bucket sort on the padded k-mer code.  reference: src/Mkvtree/ppsort.c:83-313 (bucketsortsuffixes).
 * Scanning right to left, code = code/sigma + a*sigma^(k-1); at a special (and for the sentinel) the code is
 * reset to sigma^k-1 and the next k-1 positions to the left are "special within the prefix".
 * Produces leftborder[0..numofcodes], specialcount[code] (suffixes of that bucket with a special inside the
 * k-prefix, the special-initial ones and the sentinel included) and suf[] bucket-sorted with, inside each
 * bucket, the full k-mers first.
 */
static int bucket_sort(const Text *t, uint32_t sigma, uint32_t k, uint64_t numofcodes, uint64_t *leftborder,
                       uint64_t *specialcount, uint64_t *suf)
{
  const uint64_t n = t->n, power = numofcodes / sigma;
  uint64_t *fillfull, *fillspecial;
  uint64_t code;
  int64_t inprefix;

  memset(leftborder, 0, sizeof (uint64_t) * (numofcodes + 1));
  memset(specialcount, 0, sizeof (uint64_t) * numofcodes);
  /* pass 1: counts */
  code = numofcodes - 1;
  leftborder[code]++;
  specialcount[code]++;
  inprefix = (int64_t) k - 1;
  for (uint64_t p = n; p-- > 0; )
  {
    uint8_t a = t->tis[p];
    if (IS_SPECIAL(a))
    {
      code = numofcodes - 1;
      inprefix = (int64_t) k;
    } else
    {
      code = code / sigma + (uint64_t) a * power;
    }
    leftborder[code]++;
    if (inprefix > 0)
    {
      specialcount[code]++;
      inprefix--;
    }
  }
  /* counts -> bucket starts */
  {
    uint64_t sum = 0;
    for (uint64_t c = 0; c < numofcodes; c++)
    {
      uint64_t cnt = leftborder[c];
      leftborder[c] = sum;
      sum += cnt;
    }
    leftborder[numofcodes] = sum;       /* == n+1 */
  }
  fillfull = malloc(sizeof (uint64_t) * numofcodes);
  fillspecial = malloc(sizeof (uint64_t) * numofcodes);
  if (fillfull == NULL || fillspecial == NULL)
  {
    free(fillfull);
    free(fillspecial);
    return -1;
  }
  /* pass 2: a second right-to-left scan; every part of a bucket is filled from its back
     (the order inside a part is irrelevant, it is sorted next) */
  {
    for (uint64_t c = 0; c < numofcodes; c++)
    {
      fillfull[c] = leftborder[c + 1] - specialcount[c];  /* one past the last full-k-mer slot */
      fillspecial[c] = leftborder[c + 1];                  /* one past the last special slot */
    }
    code = numofcodes - 1;
    suf[--fillspecial[code]] = n;
    inprefix = (int64_t) k - 1;
    for (uint64_t p = n; p-- > 0; )
    {
      uint8_t a = t->tis[p];
      if (IS_SPECIAL(a))
      {
        code = numofcodes - 1;
        inprefix = (int64_t) k;
      } else
      {
        code = code / sigma + (uint64_t) a * power;
      }
      if (inprefix > 0)
      {
        suf[--fillspecial[code]] = p;
        inprefix--;
      } else
      {
        suf[--fillfull[code]] = p;
      }
    }
  }
  free(fillfull);
  free(fillspecial);
  return 0;
}

/*
 * The whole path.  Any output pointer may be NULL (table not wanted), except suftab.
 *   suftab[n+1]     reference: src/Mkvtree/bese.c:855-1105 (BESESUFFIXSORT) + mkvprocess.c:329-352
 *   lcptab[n+1], llv pairs (index,value) for values >= 255, *maxbranchdepth
 *                   reference: src/Mkvtree/bese.c:533-591 (outlcpsubtab), mkvprocess.c:214-237
 *   bwttab[n+1]     reference: src/kurtz/bwtcode.c:293-311 (encodeburrowswheeler)
 *   bcktab[2*numofcodes]  reference: src/Mkvtree/mkvprocess.c:251-327 (makebcktab)
 *   stitab1[n+1]    reference: src/Mkvtree/mkvprocess.c:583-643 (makestitab1)
 *   *longest        reference: src/Mkvtree/mkvprocess.c:857-873 (determinelongest)
 * llv must have room for 2*(n+1) words if given.  prefixlength 0 = one global sort (bese.c:1048-1075).
 * Returns 0, or a negative value when out of memory / arguments are inconsistent.
 */
int oracle_esa_build(const uint8_t *tis, uint64_t n, uint32_t numofchars, uint32_t prefixlength,
                     uint64_t *suftab, uint8_t *lcptab, uint64_t *llv, uint64_t *nllv,
                     uint64_t *maxbranchdepth, uint64_t *longest, uint8_t *bwttab, uint64_t *bcktab,
                     uint8_t *stitab1)
{
  Text t = { tis, n };
  const uint64_t N = n + 1;
  uint64_t *lcpfull = NULL;

  if (suftab == NULL || numofchars == 0)
  {
    return -1;
  }
  if ((bcktab != NULL || stitab1 != NULL) && prefixlength == 0)
  {
    return -2;                                /* reference: mkvtree.c:470-473 (-bck/-sti1 require -pl) */
  }
  if (prefixlength > 0)
  {
    uint64_t numofcodes = oracle_numofcodes(numofchars, prefixlength);
    uint64_t *leftborder = malloc(sizeof (uint64_t) * (numofcodes + 1));
    uint64_t *specialcount = malloc(sizeof (uint64_t) * numofcodes);
    if (leftborder == NULL || specialcount == NULL
        || bucket_sort(&t, numofchars, prefixlength, numofcodes, leftborder, specialcount, suftab) != 0)
    {
      free(leftborder);
      free(specialcount);
      return -3;
    }
    for (uint64_t c = 0; c < numofcodes; c++)
    {
      uint64_t left = leftborder[c], right = leftborder[c + 1], mid = right - specialcount[c];
      if (mid - left > 1)
      {
        multikey_qsort(&t, suftab + left, (int64_t) (mid - left), prefixlength);  /* bese.c:1001-1006 */
      }
      if (right - mid > 1)
      {
        multikey_qsort(&t, suftab + mid, (int64_t) (right - mid), 0);             /* bese.c:1016-1020 */
      }
      if (bcktab != NULL)
      {
        bcktab[2 * c] = left;
        bcktab[2 * c + 1] = mid;
      }
    }
    free(leftborder);
    free(specialcount);
  } else
  {
    for (uint64_t i = 0; i < N; i++)
    {
      suftab[i] = i;
    }
    multikey_qsort(&t, suftab, (int64_t) N, 0);
  }
  if (longest != NULL)
  {
    for (uint64_t i = 0; i < N; i++)
    {
      if (suftab[i] == 0)
      {
        *longest = i;
        break;
      }
    }
  }
  if (lcptab != NULL || stitab1 != NULL || nllv != NULL || maxbranchdepth != NULL)
  {
    uint64_t large = 0, maxlcp = 0;
    lcpfull = malloc(sizeof (uint64_t) * N);
    if (lcpfull == NULL)
    {
      return -4;
    }
    lcpfull[0] = 0;                                                       /* bese.c:639 */
    for (uint64_t i = 1; i < N; i++)
    {
      lcpfull[i] = suffix_lcp(&t, suftab[i - 1], suftab[i], 0);
    }
    for (uint64_t i = 0; i < N; i++)
    {
      uint64_t v = lcpfull[i];
      if (v > maxlcp)
      {
        maxlcp = v;
      }
      if (lcptab != NULL)
      {
        lcptab[i] = (uint8_t) (v < 255 ? v : 255);
      }
      if (v >= 255)
      {
        if (llv != NULL)
        {
          llv[2 * large] = i;
          llv[2 * large + 1] = v;
        }
        large++;
      }
    }
    if (nllv != NULL)
    {
      *nllv = large;
    }
    if (maxbranchdepth != NULL)
    {
      *maxbranchdepth = maxlcp;
    }
  }
  if (stitab1 != NULL)
  {
    uint8_t cur = 0;
    stitab1[suftab[0]] = 0;
    for (uint64_t i = 1; i < N; i++)
    {
      uint8_t l8 = (uint8_t) (lcpfull[i] < 255 ? lcpfull[i] : 255);
      if (l8 < (uint8_t) prefixlength)
      {
        cur = 0;
      } else if (cur < 255)
      {
        cur++;
      }
      stitab1[suftab[i]] = cur;
    }
  }
  if (bwttab != NULL)
  {
    for (uint64_t i = 0; i < N; i++)
    {
      bwttab[i] = suftab[i] > 0 ? tis[suftab[i] - 1] : (uint8_t) ORACLE_UNDEFBWT;
    }
  }
  free(lcpfull);
  return 0;
}

/* Full inverse suffix array.  reference: src/Mkvtree/mksti.c:15-34 */
void oracle_inverse_suftab(const uint64_t *suftab, uint64_t n, uint64_t *sti)
{
  for (uint64_t i = 0; i <= n; i++)
  {
    sti[suftab[i]] = i;
  }
}

/*
 * Skip table.  reference: src/kurtz/mkskip.c:53-91 (makeskiptable): a stack of (depth, step) pairs walks the lcp
 * table; an entry is popped by the first later lcp value that is strictly smaller and gets the index in front of
 * it, entries that survive get totallength.  lcp values of 255 are read from the llv pairs in order.
 * skiptable has n + 1 entries.  Returns 0, or -1 when memory is exhausted.
 */
int oracle_skiptable(const uint8_t *lcptab, const uint64_t *llv, uint64_t n, uint64_t *skiptable)
{
  uint64_t *depth = (uint64_t *) malloc((n + 2) * sizeof(uint64_t));
  uint64_t *step = (uint64_t *) malloc((n + 2) * sizeof(uint64_t));
  uint64_t top = 0, exception = 0;
  if (depth == NULL || step == NULL)
  {
    free(depth);
    free(step);
    return -1;
  }
  depth[top] = 0;
  step[top] = 0;
  top++;
  for (uint64_t i = 1; i <= n; i++)
  {
    uint64_t cur = lcptab[i];
    if (cur == 255)
    {
      cur = llv[2 * exception + 1];
      exception++;
    }
    while (cur < depth[top - 1])
    {
      skiptable[step[top - 1]] = i - 1;
      top--;
    }
    depth[top] = cur;
    step[top] = i;
    top++;
  }
  while (top > 0)
  {
    skiptable[step[top - 1]] = n;
    top--;
  }
  free(depth);
  free(step);
  return 0;
}

/*
 * Child table.  reference: src/kurtz/mkcld.c:41-108 (mkcldtabnextlIndex), :110-207 (mkcldtabupdown), :225-284
 * (compresscldtab); driver src/Mkvtree/mkcld.mn.c:13-47.  Two walks over the lcp table with a stack of (depth, step)
 * pairs that starts with (0, 0):
 *   nextlIndex: pop everything deeper than the current value; if the top then has the same depth, the top's entry
 *               points to the current index.  Push.
 *   up / down:  while popping, remember the entry popped last; whenever the new top is at least as deep as the current
 *               value and differs in depth from the entry just popped, the top's `down` points to that entry.  After the pops the current index gets `up` = distance to the entry popped last
 *               (once).  Push.
 * All three fields are distances, stored in one byte: 0 = undefined, 255 = "255 or more" (virtualdef.h:165-166).
 * cld receives n + 1 records {up, down, nextlIndex} (the .cld file), cld1 -- when not NULL -- the n + 1 bytes of the
 * .cld1 file: three loops over the same byte, the later one overwrites, none touches index n (and the third stops at
 * n - 2); bytes no loop writes are 0 here (the reference leaves what malloc returned).
 * Returns 0, or -1 when memory is exhausted.
 */
int oracle_childtab(const uint8_t *lcptab, const uint64_t *llv, uint64_t n, uint8_t *cld, uint8_t *cld1)
{
  uint64_t *depth = (uint64_t *) malloc((n + 2) * sizeof(uint64_t));
  uint64_t *step = (uint64_t *) malloc((n + 2) * sizeof(uint64_t));
  uint64_t *lcp = (uint64_t *) malloc((n + 1) * sizeof(uint64_t));
  uint64_t top, exception = 0;
  if (depth == NULL || step == NULL || lcp == NULL)
  {
    free(depth);
    free(step);
    free(lcp);
    return -1;
  }
  lcp[0] = lcptab[0];
  for (uint64_t i = 1; i <= n; i++)
  {
    lcp[i] = lcptab[i];
    if (lcp[i] == 255)
    {
      lcp[i] = llv[2 * exception + 1];
      exception++;
    }
  }
  memset(cld, 0, 3 * (n + 1));
#define CLD_CLAMP(D) ((uint8_t) ((D) >= 255 ? 255 : (D)))
  /* walk 1: nextlIndex */
  top = 0;
  depth[top] = 0;
  step[top] = 0;
  top++;
  for (uint64_t i = 1; i <= n; i++)
  {
    while (lcp[i] < depth[top - 1])
    {
      top--;
    }
    if (lcp[i] == depth[top - 1])
    {
      cld[3 * step[top - 1] + 2] = CLD_CLAMP(i - step[top - 1]);
    }
    depth[top] = lcp[i];
    step[top] = i;
    top++;
  }
  /* walk 2: up and down */
  top = 0;
  depth[top] = 0;
  step[top] = 0;
  top++;
  {
    uint64_t last_step = 0, last_depth = 0;  /* step 0 = nothing popped since the last `up` (index 0 is never popped) */
    for (uint64_t i = 1; i <= n; i++)
    {
      while (lcp[i] < depth[top - 1])
      {
        last_step = step[top - 1];
        last_depth = depth[top - 1];
        top--;
        if (lcp[i] <= depth[top - 1] && depth[top - 1] != last_depth)
        {
          cld[3 * step[top - 1] + 1] = CLD_CLAMP(last_step - step[top - 1]);
        }
      }
      if (last_step != 0)
      {
        cld[3 * i] = CLD_CLAMP(i - last_step);
        last_step = 0;
      }
      depth[top] = lcp[i];
      step[top] = i;
      top++;
    }
  }
  if (cld1 != NULL)
  {
    memset(cld1, 0, n + 1);
    for (uint64_t i = 0; i < n; i++)
    {
      if (cld[3 * i + 2] > 0)
      {
        cld1[i] = cld[3 * i + 2];
      }
    }
    for (uint64_t i = 0; i < n; i++)
    {
      uint64_t next = i + cld[3 * i + 2], down = i + cld[3 * i + 1];
      if ((next == i && down > i) || next == n)
      {
        cld1[i] = CLD_CLAMP(down - i);
      }
    }
    for (uint64_t i = 0; i + 1 < n; i++)
    {
      if (lcp[i] > lcp[i + 1])
      {
        uint64_t up = i + 1 - cld[3 * (i + 1)], down = i + 1 + cld[3 * (i + 1) + 1];
        if (up != down)
        {
          cld1[i] = CLD_CLAMP(i + 1 - up);
        }
      }
    }
  }
#undef CLD_CLAMP
  free(depth);
  free(step);
  free(lcp);
  return 0;
}

/*
 * Property checks restating the reference's DEBUG invariants; return 0 when the tables are consistent,
 * otherwise a positive code naming the first violated property.
 *   1: suftab is not a permutation of 0..n            reference: src/Mkvtree/bese.c:355-451 (checksuftab)
 *   2: two adjacent suffixes are out of order
 *   3: an lcp value (with llv overflow) is wrong      reference: src/Mkvtree/bese.c:454-528 (checklcpsubtab)
 *   4: bwt does not decode to the text                reference: src/Mkvtree/cprsuf.c:66-95 (checksufencoding)
 *   5: bck.left + sti1 != inverse suftab (saturated)  reference: src/Mkvtree/mkvprocess.c:672-763 (checkstitab1)
 * Tables passed as NULL are skipped.
 */
int oracle_check_tables(const uint8_t *tis, uint64_t n, uint32_t numofchars, uint32_t prefixlength,
                        const uint64_t *suftab, const uint8_t *lcptab, const uint64_t *llv, uint64_t nllv,
                        const uint8_t *bwttab, const uint64_t *bcktab, const uint8_t *stitab1)
{
  Text t = { tis, n };
  const uint64_t N = n + 1;
  uint8_t *seen = calloc(N, 1);
  if (seen == NULL)
  {
    return -1;
  }
  for (uint64_t i = 0; i < N; i++)
  {
    if (suftab[i] > n || seen[suftab[i]])
    {
      free(seen);
      return 1;
    }
    seen[suftab[i]] = 1;
  }
  free(seen);
  {
    uint64_t nextlarge = 0;
    for (uint64_t i = 1; i < N; i++)
    {
      uint64_t l = suffix_lcp(&t, suftab[i - 1], suftab[i], 0);
      if (sym(&t, suftab[i - 1] + l) > sym(&t, suftab[i] + l))
      {
        return 2;
      }
      if (lcptab != NULL)
      {
        if (l < 255)
        {
          if (lcptab[i] != l)
          {
            return 3;
          }
        } else
        {
          if (lcptab[i] != 255)
          {
            return 3;
          }
          if (llv != NULL)
          {
            if (nextlarge >= nllv || llv[2 * nextlarge] != i || llv[2 * nextlarge + 1] != l)
            {
              return 3;
            }
            nextlarge++;
          }
        }
      }
    }
    if (lcptab != NULL && (lcptab[0] != 0 || (llv != NULL && nextlarge != nllv)))
    {
      return 3;
    }
  }
  if (bwttab != NULL)
  {
    for (uint64_t i = 0; i < N; i++)
    {
      uint8_t expect = suftab[i] > 0 ? tis[suftab[i] - 1] : (uint8_t) ORACLE_UNDEFBWT;
      if (bwttab[i] != expect)
      {
        return 4;
      }
    }
  }
  if (bcktab != NULL && stitab1 != NULL && prefixlength > 0)
  {
    uint64_t numofcodes = oracle_numofcodes(numofchars, prefixlength);
    for (uint64_t c = 0; c < numofcodes; c++)
    {
      for (uint64_t i = bcktab[2 * c]; i < bcktab[2 * c + 1]; i++)
      {
        uint64_t off = i - bcktab[2 * c];
        if (stitab1[suftab[i]] != (off < 255 ? off : 255))
        {
          return 5;
        }
      }
    }
  }
  return 0;
}
