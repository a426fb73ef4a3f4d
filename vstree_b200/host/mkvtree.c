/*
 * mkvtree.c -- host driver of the B200-native mkvtree: the reference's command line and on-disk index layout,
 * with the enhanced-suffix-array construction done by libvstree_b200 (CUDA, sm_100a) through its C ABI.
 *
 * Written from scratch; it restates the BEHAVIOUR of
 *   option parsing        src/Mkvtree/mkvtree.c:169-551 (mkvparseoptions), src/kurtz-basic/procopt.c:103-129
 *   input and alphabets   src/Mkvtree/mkvinput.c, src/kurtz-basic/multiseq-adv.c:1036-1291,1360-1453,
 *                         src/kurtz-basic/alphabet.c:180-410, src/kurtz-basic/parsemultiform.c:352-389
 *   table writers         src/Mkvtree/mkvprocess.c:38-211,403-580,875-1089
 *   main                  src/Mkvtree/mkvfile.c:37-79
 * The host keeps alphabet transformation and file I/O; there is no CPU implementation of the sort behind it:
 * without a GPU the program fails.
 *
 *   -rev / -cpl           src/Mkvtree/mkvinput.c:273-309, src/kurtz/reverse.c:9-70, src/Mkvtree/mkvtree.c:143-161
 *
 *   GenBank/EMBL/SwissProt  src/kurtz-basic/parsemultiform.c:225-385 (single input file)
 *
 *   no alphabet option      src/Mkvtree/mkvinput.c:100-170, findalphabet src/kurtz-basic/alphabet.c:456-487, .al2
 *
 * Not supported (reported as an error): -maxdepth.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <limits.h>
#include <pthread.h>
#include <sched.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "vstree_b200.h"

#define AUTOPREFIXLENGTH 1024u          /* src/Mkvtree/mkvtree.c (AUTOPREFIXLENGTH) */
#define MAXNUMBEROFFILES 256u
#define UNDEFCHAR 253u                  /* src/include/chardef.h */
#define WILDCARD ESA_WILDCARD
#define SEPARATOR ESA_SEPARATOR
#define ISSPECIAL(c) ((c) >= WILDCARD)
#define OUTPUTTABS (ESA_TISTAB | ESA_OISTAB | ESA_BWTTAB | ESA_SUFTAB | ESA_LCPTAB | ESA_BCKTAB | ESA_STI1TAB | ESA_SKPTAB)

static char message[4096 + 2 * PATH_MAX];

static int fail(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(message, sizeof message, fmt, ap);
  va_end(ap);
  return -1;
}

static int verbose = 0;
static void say(const char *fmt, ...)
{
  if (verbose)
  {
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
    printf("\n");
  }
}

/* ------------------------------------------------------------------------------------------------------ */
/* options */

enum
{
  OPTDB, OPTQ, OPTSMAP, OPTDNA, OPTPROTEIN, OPTTRANSNUM, OPTPL, OPTREV, OPTCPL, OPTINDEXNAME, OPTTIS, OPTOIS, OPTSUF,
  OPTSTI1, OPTBWT, OPTBCK, OPTLCP, OPTSKP, OPTALLOUT, OPTMAXDEPTH, OPTV, OPTVERSION, OPTHELP, NUMOFOPTIONS
};

static const char *optname[NUMOFOPTIONS] = {
  "-db", "-q", "-smap", "-dna", "-protein", "-transnum", "-pl", "-rev", "-cpl", "-indexname", "-tis", "-ois", "-suf",
  "-sti1", "-bwt", "-bck", "-lcp", "-skp", "-allout", "-maxdepth", "-v", "-version", "-help"
};

static const char *optdesc[NUMOFOPTIONS] = {
  "specify database files (mandatory)", "specify query files", "specify file containing a symbol mapping",
  "input is DNA sequence", "input is Protein sequence", "(only for mkdna6idx)",
  "specify prefix length for bucket sort\nrecommendation: use without argument;\nthen a reasonable prefix length is "
  "automatically determined",
  "use reverse input sequence", "use reverse complement of the input sequence",
  "specify name for index to be generated", "output transformed input sequences (tistab) to file",
  "output original input sequences (oistab) to file", "output suffix array (suftab) to file",
  "output reduced inverse suffix array (sti1tab) to file", "output Burrows-Wheeler Transformation (bwttab) to file",
  "output bucket boundaries (bcktab) to file", "output longest common prefix lengths (lcptab) to file",
  "output skip values (skptab) to file", "output all index tables to files",
  "restrict the sorting to prefixes of the given length", "verbose mode", "show the version of the program",
  "this option"
};

typedef struct
{
  int isset[NUMOFOPTIONS];
  uint32_t demand;
  uint32_t prefixlength;
  const char *smapfile;
  const char *files[MAXNUMBEROFFILES];   /* database files, then query files (option -q) */
  unsigned numoffiles;                   /* all of them */
  unsigned numofqueryfiles;
  const char *queryfiles[MAXNUMBEROFFILES];
  char indexname[PATH_MAX + 1];
} Options;

static int isoption(const char *s) { return s[0] == '-' && isalpha((unsigned char) s[1]); }

static int imply(const Options *o, int a, int b)
{
  if (o->isset[a] && !o->isset[b])
  {
    return fail("option %s requires option %s", optname[a], optname[b]);
  }
  return 0;
}

static int exclude(const Options *o, int a, int b)
{
  if (o->isset[a] && o->isset[b])
  {
    return fail("option %s and option %s exclude each other", optname[a], optname[b]);
  }
  return 0;
}

static void showhelp(const char *prog)
{
  printf("%s: options\n", prog);
  for (int i = 0; i < NUMOFOPTIONS; i++)
  {
    if (i == OPTTRANSNUM)
    {
      continue;
    }
    const char *d = optdesc[i];
    printf("%-11s", optname[i]);
    for (; *d; d++)
    {
      putchar(*d);
      if (*d == '\n')
      {
        printf("%-11s", "");
      }
    }
    putchar('\n');
  }
}

/* returns 0 ok, 1 help/version shown, -1 usage, -2 other error */
static int parseoptions(Options *o, int argc, char **argv)
{
  int argnum;
  memset(o, 0, sizeof *o);
  if (argc == 1)
  {
    fail("Usage: %s options indexname\n%s -help shows possible options", argv[0], argv[0]);
    return -1;
  }
  for (argnum = 1; argnum < argc && isoption(argv[argnum]); argnum++)
  {
    int opt = -1;
    long readint;
    if (strcmp("-prefix", argv[argnum]) == 0)
    {
      fprintf(stderr, "%s: option -prefix has been renamed to %s", argv[0], optname[OPTINDEXNAME]);
      exit(EXIT_FAILURE);
    }
    for (int i = 0; i < NUMOFOPTIONS; i++)
    {
      if (strcmp(argv[argnum], optname[i]) == 0)
      {
        opt = i;
        break;
      }
    }
    if (opt < 0)
    {
      fail("illegal option \"%s\"", argv[argnum]);
      return -2;
    }
    if (o->isset[opt])
    {
      fail("option \"%s\" already set", argv[argnum]);
      return -2;
    }
    o->isset[opt] = 1;
    switch (opt)
    {
      case OPTHELP:
        showhelp(argv[0]);
        return 1;
      case OPTVERSION:
        printf("mkvtree (Vmatch-compatible index construction) %s\n", esa_version());
        return 1;
      case OPTSMAP:
      case OPTINDEXNAME:
        argnum++;
        if (argnum >= argc || isoption(argv[argnum]))
        {
          fail("missing argument to option %s", optname[opt]);
          return -2;
        }
        if (opt == OPTSMAP)
        {
          o->smapfile = argv[argnum];
        } else
        {
          if (strlen(argv[argnum]) > PATH_MAX)
          {
            fail("string \"%s\" is too long, cannot copy", argv[argnum]);
            return -2;
          }
          strcpy(o->indexname, argv[argnum]);
        }
        break;
      case OPTTRANSNUM:
        fail("illegal option \"%s\"", optname[OPTTRANSNUM]);
        return -2;
      case OPTPL:
        if (argnum + 1 < argc && !isoption(argv[argnum + 1]))
        {
          argnum++;
          if (sscanf(argv[argnum], "%ld", &readint) != 1 || readint <= 0)
          {
            fail("argument \"%s\" of option %s must be positive number", argv[argnum], optname[OPTPL]);
            return -2;
          }
          o->prefixlength = (uint32_t) readint;
          if (o->prefixlength == AUTOPREFIXLENGTH)
          {
            fail("Illegal value %lu as argument to option %s", (unsigned long) o->prefixlength, optname[OPTPL]);
            return -2;
          }
        } else
        {
          o->prefixlength = AUTOPREFIXLENGTH;
        }
        break;
      case OPTTIS: o->demand |= ESA_TISTAB; break;
      case OPTOIS: o->demand |= ESA_OISTAB; break;
      case OPTSUF: o->demand |= ESA_SUFTAB; break;
      case OPTSTI1: o->demand |= ESA_STI1TAB | ESA_LCPTAB | ESA_BCKTAB; break;
      case OPTBWT: o->demand |= ESA_BWTTAB; break;
      case OPTBCK: o->demand |= ESA_BCKTAB; break;
      case OPTLCP: o->demand |= ESA_LCPTAB; break;
      case OPTSKP: o->demand |= ESA_SKPTAB; break;
      case OPTV: verbose = 1; break;
      case OPTMAXDEPTH:
        if (!(argnum == argc - 1 || isoption(argv[argnum + 1])))
        {
          argnum++;
          if (sscanf(argv[argnum], "%ld", &readint) != 1 || readint <= 0)
          {
            fail("argument \"%s\" of option %s must be positive number", argv[argnum], optname[OPTMAXDEPTH]);
            return -2;
          }
        }
        break;
      case OPTDB:
      case OPTQ:
        argnum++;
        while (argnum < argc)
        {
          if (isoption(argv[argnum]))
          {
            argnum--;
            break;
          }
          if (opt == OPTDB && o->numoffiles < MAXNUMBEROFFILES)
          {
            o->files[o->numoffiles] = argv[argnum];
          }
          if (opt == OPTDB)
          {
            o->numoffiles++;
          }
          if (opt == OPTQ && o->numofqueryfiles < MAXNUMBEROFFILES)
          {
            o->queryfiles[o->numofqueryfiles] = argv[argnum];
          }
          if (opt == OPTQ)
          {
            o->numofqueryfiles++;
          }
          argnum++;
        }
        break;
      default:
        break;
    }
  }
  if (argc > argnum)
  {
    fail("illegal additional argument \"%s\"", argv[argc - 1]);
    return -2;
  }
  if (imply(o, OPTALLOUT, OPTPL) || imply(o, OPTSTI1, OPTPL) || imply(o, OPTBCK, OPTPL) || imply(o, OPTSKP, OPTLCP)
      || imply(o, OPTSKP, OPTSUF) || exclude(o, OPTDNA, OPTSMAP) || exclude(o, OPTPROTEIN, OPTSMAP)
      || exclude(o, OPTDNA, OPTPROTEIN))
  {
    return -2;
  }
  if (o->isset[OPTALLOUT])
  {
    if (o->demand & OUTPUTTABS)
    {
      fail("option %s cannot be combined with any other output option", optname[OPTALLOUT]);
      return -2;
    }
    o->demand = OUTPUTTABS;
  }
  if (o->numoffiles == 0)
  {
    fail("option %s is mandatory", optname[OPTDB]);
    return -2;
  }
  if (o->numoffiles + o->numofqueryfiles >= MAXNUMBEROFFILES)
  {
    fail("too many input files, maximum is %lu", (unsigned long) MAXNUMBEROFFILES);
    return -2;
  }
  /* the query files follow the database files (src/Mkvtree/mkvtree.c:502-515); from here on they are input files
     like the others, only the project file tells them apart */
  for (unsigned i = 0; i < o->numofqueryfiles; i++)
  {
    o->files[o->numoffiles++] = o->queryfiles[i];
  }
  /* what this build does not do (see the header comment) */
  {
    static const int unsupported[] = {OPTMAXDEPTH};
    for (size_t i = 0; i < sizeof unsupported / sizeof unsupported[0]; i++)
    {
      if (o->isset[unsupported[i]])
      {
        fail("option %s is not supported by the GPU build of mkvtree", optname[unsupported[i]]);
        return -2;
      }
    }
  }
  if (o->isset[OPTINDEXNAME])
  {
    if (!(o->demand & OUTPUTTABS))
    {
      fail("option %s only allowed in connection with output option", optname[OPTINDEXNAME]);
      return -2;
    }
  } else if (o->demand & OUTPUTTABS)
  {
    const char *base, *slash;
    if (o->numoffiles > 1)
    {
      fail("option %s required for multiple input files", optname[OPTINDEXNAME]);
      return -2;
    }
    slash = strrchr(o->files[0], '/');
    base = slash != NULL && slash[1] != '\0' ? slash + 1 : o->files[0];
    if (strlen(base) > PATH_MAX)
    {
      fail("string \"%s\" is too long, cannot copy", base);
      return -2;
    }
    strcpy(o->indexname, base);
  }
  if ((o->demand & OUTPUTTABS) && (o->isset[OPTREV] || o->isset[OPTCPL]))
  {
    /* addindexsuffix, src/Mkvtree/mkvtree.c:143-161: the index of a transformed sequence gets its own name */
    const char *suffix = o->isset[OPTREV] ? (o->isset[OPTCPL] ? ".rcp" : ".rev") : ".cpl";
    if (strlen(o->indexname) + strlen(suffix) > PATH_MAX)
    {
      fail("string \"%s\" is too long, cannot copy", o->indexname);
      return -2;
    }
    strcat(o->indexname, suffix);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* -rev / -cpl: the TRANSFORMED sequence (and nothing else: .ois, .des, .sds, .ssp and the counts of the project file
   describe the sequence as it was read) is reversed and/or complemented in place, separators and wildcards included
   (src/Mkvtree/mkvinput.c:273-309, src/kurtz/reverse.c:9-70).  Complement: code c of 0..3 becomes 3 - c, specials
   stay; any other code is an error. */
static int complementcode(uint8_t cc, uint8_t *out)
{
  if (ISSPECIAL(cc))
  {
    *out = cc;
    return 0;
  }
  if (cc > 3)
  {
    return fail("complement of %lu undefined", (unsigned long) cc);
  }
  *out = (uint8_t) (3 - cc);
  return 0;
}

static int transformsequence(uint8_t *seq, uint64_t len, int rev, int cpl)
{
  if (len == 0)
  {
    return 0;
  }
  if (rev && cpl)
  {
    uint8_t *front = seq, *back = seq + len - 1;
    say("reverse complement sequence");
    for (; front <= back; front++, back--)
    {
      uint8_t tmp = 0, other = 0;
      if (complementcode(*front, &tmp) != 0 || complementcode(*back, &other) != 0)
      {
        return -1;
      }
      *front = other;
      *back = tmp;
    }
  } else if (rev)
  {
    uint8_t *front = seq, *back = seq + len - 1;
    say("reverse sequence");
    for (; front < back; front++, back--)
    {
      const uint8_t tmp = *front;
      *front = *back;
      *back = tmp;
    }
  } else if (cpl)
  {
    say("complement sequence");
    for (uint64_t i = 0; i < len; i++)
    {
      if (complementcode(seq[i], &seq[i]) != 0)
      {
        return -1;
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* alphabets */

typedef struct
{
  unsigned symbolmap[UCHAR_MAX + 1];  /* UNDEFCHAR = not in the alphabet */
  unsigned char mapdomain[UCHAR_MAX + 1];
  unsigned char characters[UCHAR_MAX + 1];
  unsigned domainsize, mapsize;
  int fromtext;  /* no -dna/-protein/-smap: the alphabet is the set of bytes of the (plain) input, findalphabet,
                    src/kurtz-basic/alphabet.c:456-487; there are no wildcards then, mapsize = alphabet size */
} Alphabet;

static void assign_builtin(Alphabet *a, const char *letters, int pairs, const char *wildcards, const char *domain)
{
  for (unsigned i = 0; i <= UCHAR_MAX; i++)
  {
    a->symbolmap[i] = UNDEFCHAR;
  }
  unsigned code = 0;
  for (const char *p = letters; *p; code++)
  {
    /* DNA: "aA","cC","gG","tTuU" groups are given explicitly by the caller through `pairs` */
    if (pairs)
    {
      while (*p && *p != ' ')
      {
        a->symbolmap[(unsigned char) *p++] = code;
      }
      if (*p == ' ')
      {
        p++;
      }
    } else
    {
      a->symbolmap[(unsigned char) *p++] = code;
    }
  }
  for (const char *p = wildcards; *p; p++)
  {
    a->symbolmap[(unsigned char) *p] = WILDCARD;
  }
  a->domainsize = (unsigned) strlen(domain);
  memcpy(a->mapdomain, domain, a->domainsize);
  a->mapsize = code + 1;
}

static FILE *scanpathsforfile(const char *envname, const char *filename)
{
  FILE *fp = fopen(filename, "rb");
  const char *env;
  if (fp != NULL)
  {
    return fp;
  }
  if (strchr(filename, '/') != NULL)
  {
    fail("filename \"%s\" contains illegal symbol '%c': the path list specified by environment variable \"%s\" "
         "cannot be searched for it", filename, '/', envname);
    return NULL;
  }
  env = getenv(envname);
  if (env != NULL)
  {
    char *copy = strdup(env), *save = NULL;
    for (char *dir = strtok_r(copy, ":", &save); dir != NULL; dir = strtok_r(NULL, ":", &save))
    {
      char path[2 * PATH_MAX + 2];
      snprintf(path, sizeof path, "%s/%s", dir, filename);
      fp = fopen(path, "rb");
      if (fp != NULL)
      {
        free(copy);
        return fp;
      }
    }
    free(copy);
  }
  fail("cannot find file \"%s\"", filename);
  return NULL;
}

/* symbol map file: one line per code, the last line holds the wildcards; "xyz c" shows the class as c
   (src/kurtz-basic/alphabet.c:180-290) */
static int read_symbolmap(Alphabet *a, const char *mapfile)
{
  FILE *fp = scanpathsforfile("MKVTREESMAPDIR", mapfile);
  char *line = NULL;
  size_t cap = 0;
  ssize_t len;
  unsigned long linecount = 0;
  int preamble = 1;
  if (fp == NULL)
  {
    return -1;
  }
  a->domainsize = a->mapsize = 0;
  for (unsigned i = 0; i <= UCHAR_MAX; i++)
  {
    a->symbolmap[i] = UNDEFCHAR;
  }
  while ((len = getline(&line, &cap, fp)) != -1)
  {
    ssize_t i;
    int blankfound = 0;
    if (len > 0 && line[len - 1] == '\n')
    {
      line[--len] = '\0';
    }
    linecount++;
    if (len == 0)
    {
      continue;
    }
    if (preamble)
    {
      if (line[0] == '#')
      {
        continue;
      }
      preamble = 0;
    }
    for (i = 0; i < len; i++)
    {
      unsigned char cc = (unsigned char) line[i];
      if (ispunct(cc) || isalnum(cc))
      {
        if (a->symbolmap[cc] != UNDEFCHAR)
        {
          fclose(fp);
          free(line);
          return fail("cannot map symbol '%c' to %lu: it is already mapped to %lu", cc, (unsigned long) a->mapsize,
                      (unsigned long) a->symbolmap[cc]);
        }
        a->symbolmap[cc] = a->mapsize;
        a->mapdomain[a->domainsize++] = cc;
      } else if (cc == ' ')
      {
        blankfound = 1;
        break;
      } else
      {
        fclose(fp);
        free(line);
        return fail("illegal character '%c' in line %lu of mapfile %s", cc, linecount, mapfile);
      }
    }
    if (blankfound)
    {
      if (isspace((unsigned char) line[i + 1]) || line[i + 1] == '\0')
      {
        fclose(fp);
        free(line);
        return fail("illegal character '%c' at the end of line %lu in mapfile %s", line[i + 1], linecount, mapfile);
      }
      a->characters[a->mapsize++] = (unsigned char) line[i + 1];
    } else
    {
      a->characters[a->mapsize++] = (unsigned char) line[0];
    }
  }
  free(line);
  fclose(fp);
  if (a->mapsize == 0)
  {
    return fail("symbol map file \"%s\" defines no symbols", mapfile);
  }
  for (unsigned i = 0; i <= UCHAR_MAX; i++)
  {
    if (a->symbolmap[i] == a->mapsize - 1)
    {
      a->symbolmap[i] = WILDCARD;
    }
  }
  a->characters[WILDCARD] = a->characters[a->mapsize - 1];
  return 0;
}

static int make_alphabet(Alphabet *a, const Options *o)
{
  memset(a, 0, sizeof *a);
  if (o->isset[OPTDNA])
  {
    assign_builtin(a, "aA cC gG tTuU", 1, "nsywrkvbdhmNSYWRKVBDHM", "aAcCgGtTuUnsywrkvbdhmNSYWRKVBDHM");
    memcpy(a->characters, "acgtn", 5);
    a->characters[WILDCARD] = 'n';
    return 0;
  }
  if (o->isset[OPTPROTEIN])
  {
    assign_builtin(a, "LVIFKREDAGSTNQYWPHMC", 0, "XUBZJO*-", "LVIFKREDAGSTNQYWPHMCXUBZJO*-");
    memcpy(a->characters, "LVIFKREDAGSTNQYWPHMCX", 21);
    a->characters[WILDCARD] = 'X';
    return 0;
  }
  if (o->isset[OPTSMAP])
  {
    return read_symbolmap(a, o->smapfile);
  }
  a->fromtext = 1;  /* plain input only; decided when the input has been read (mkvinput.c:100-170) */
  return 0;
}

static void verbosealphabet(const Alphabet *a)
{
  if (!verbose)
  {
    return;
  }
  if (a->fromtext)
  {
    /* the reference prints this alphabet through the arrays of a symbol-map alphabet it never filled: the domain is
       all zero bytes, and the place of the last character shows the (unset) wildcard character */
    printf("alphabet \"");
    for (unsigned i = 0; i < a->mapsize; i++)
    {
      printf("\\0");
    }
    printf("\" (size %lu) mapped to \"", (unsigned long) a->mapsize);
    for (unsigned i = 0; i + 1 < a->mapsize; i++)
    {
      unsigned char c = a->characters[i];
      if (isprint(c) && !isspace(c))
      {
        putchar(c);
      } else
      {
        printf("\\%lu", (unsigned long) c);
      }
    }
    printf("\\0\" (size %lu)\n", (unsigned long) a->mapsize);
    return;
  }
  printf("alphabet \"");
  for (unsigned i = 0; i < a->domainsize; i++)
  {
    unsigned char c = a->mapdomain[i];
    if (isprint(c))
    {
      putchar(c);
    } else
    {
      printf("\\%lu", (unsigned long) c);
    }
  }
  printf("\" (size %lu) mapped to \"", (unsigned long) a->domainsize);
  for (unsigned i = 0; i < a->mapsize; i++)
  {
    unsigned char c = a->characters[i];
    if (isprint(c))
    {
      putchar(c);
    } else
    {
      printf("\\%lu", (unsigned long) c);
    }
  }
  printf("\" (size %lu)\n", (unsigned long) a->mapsize);
}

/* ------------------------------------------------------------------------------------------------------ */
/* input */

typedef struct
{
  uint8_t *tis, *ois;            /* transformed / original sequence, totallength bytes */
  uint8_t *sortseq;              /* what the sorter gets when that is not tis (raw text without an alphabet) */
  uint64_t totallength;
  uint8_t *desc;                 /* descriptions, each ending in '\n' */
  uint64_t desclen;
  uint64_t *startdesc;           /* numofsequences + 1 offsets into desc */
  uint64_t *markpos;             /* numofsequences - 1 separator positions */
  uint64_t numofsequences;
  uint64_t specialcharacters, specialranges, lengthofspecialprefix, lengthofspecialsuffix;
  uint64_t filelength[MAXNUMBEROFFILES];
  uint64_t filesep[MAXNUMBEROFFILES];   /* separator position in tis between file i and i+1 */
} Multiseq;

static int hasgzsuffix(const char *name)
{
  size_t l = strlen(name);
  return l > 3 && strcmp(name + l - 3, ".gz") == 0;
}

static uint8_t *readwholefile(const char *name, uint64_t *len)
{
  uint8_t *buf = NULL;
  uint64_t cap = 0, used = 0;
  if (hasgzsuffix(name))
  {
    gzFile gz = gzopen(name, "rb");
    if (gz == NULL)
    {
      fail("cannot read file \"%s\": open failed: %s", name, strerror(errno));
      return NULL;
    }
    for (;;)
    {
      int got;
      if (cap - used < (1u << 20))
      {
        cap = cap ? 2 * cap : (1u << 22);
        buf = realloc(buf, cap + 1);
        if (buf == NULL)
        {
          fail("not enough memory to read \"%s\"", name);
          return NULL;
        }
      }
      got = gzread(gz, buf + used, (unsigned) (cap - used > (1u << 30) ? (1u << 30) : cap - used));
      if (got < 0)
      {
        fail("cannot read gzipped file \"%s\"", name);
        free(buf);
        gzclose(gz);
        return NULL;
      }
      if (got == 0)
      {
        break;
      }
      used += (uint64_t) got;
    }
    gzclose(gz);
  } else
  {
    FILE *fp = fopen(name, "rb");
    struct stat st;
    if (fp == NULL || fstat(fileno(fp), &st) != 0)
    {
      fail("cannot read file \"%s\": open failed: %s", name, strerror(errno));
      if (fp != NULL)
      {
        fclose(fp);
      }
      return NULL;
    }
    cap = (uint64_t) st.st_size;
    buf = malloc(cap + 1);
    if (buf == NULL)
    {
      fclose(fp);
      fail("not enough memory to read \"%s\"", name);
      return NULL;
    }
    used = fread(buf, 1, cap, fp);
    fclose(fp);
    if (used != cap)
    {
      free(buf);
      fail("cannot read file \"%s\": short read", name);
      return NULL;
    }
  }
  buf[used] = 0;
  *len = used;
  return buf;
}

/* several files -> one buffer; a file that does not start with '>' gets a ">filename" line
   (src/kurtz-basic/multiseq-adv.c:1360-1453) */
static uint8_t *concatfiles(const Options *o, Multiseq *ms, uint64_t *concatfilesep, uint64_t *totalbytes)
{
  uint8_t *concat = NULL;
  uint64_t used = 0, cap = 0;
  for (unsigned f = 0; f < o->numoffiles; f++)
  {
    uint64_t len, skip = 0, namelen = strlen(o->files[f]);
    uint8_t *buf = readwholefile(o->files[f], &len);
    int isfasta = 0;
    if (buf == NULL)
    {
      free(concat);
      return NULL;
    }
    ms->filelength[f] = len;
    for (uint64_t i = 0; i < len; i++)
    {
      if (isspace(buf[i]))
      {
        skip++;
      } else
      {
        isfasta = buf[i] == '>';
        break;
      }
    }
    if (used + len + namelen + 8 > cap)
    {
      cap = 2 * (used + len + namelen + 8);
      concat = realloc(concat, cap);
      if (concat == NULL)
      {
        free(buf);
        fail("not enough memory for the concatenated input");
        return NULL;
      }
    }
    if (f > 0)
    {
      concatfilesep[f - 1] = used;
    }
    if (!isfasta)
    {
      uint64_t show = hasgzsuffix(o->files[f]) ? namelen - 3 : namelen;
      printf(">%.*s\n", (int) show, o->files[f]);  /* the reference prints this line, verbose or not */
      used += (uint64_t) sprintf((char *) concat + used, ">%.*s\n", (int) show, o->files[f]);
    }
    say("reading file \"%s\"", o->files[f]);
    memcpy(concat + used, buf + skip, len - skip);
    used += len - skip;
    if (!isfasta)
    {
      concat[used++] = '\n';
    }
    free(buf);
  }
  *totalbytes = used;
  return concat;
}

#define PUSH(ARR, CNT, CAP, VAL)                                   \
  do                                                               \
  {                                                                \
    if ((CNT) == (CAP))                                            \
    {                                                              \
      (CAP) = (CAP) ? 2 * (CAP) : 1024;                            \
      (ARR) = realloc((ARR), (CAP) * sizeof *(ARR));               \
      if ((ARR) == NULL)                                           \
      {                                                            \
        return fail("not enough memory while parsing the input");  \
      }                                                            \
    }                                                              \
    (ARR)[(CNT)++] = (VAL);                                        \
  } while (0)

/* what follows the sweep over the input: special prefix and suffix, empty sequences (both parsers end here) */
static int finishfasta(Multiseq *ms)
{
  const uint64_t out = ms->totallength;
  if (ms->numofsequences == 0)
  {
    return fail("no sequences in multiple fasta file");
  }
  for (uint64_t p = 0; p < out && ISSPECIAL(ms->tis[p]); p++)
  {
    ms->lengthofspecialprefix++;
  }
  if (out > 0)
  {
    for (uint64_t p = out - 1; ISSPECIAL(ms->tis[p]); p--)
    {
      ms->lengthofspecialsuffix++;
      if (p == 0)
      {
        break;
      }
    }
  }
  if (ms->numofsequences >= 2)
  {
    for (uint64_t m = 1; m < ms->numofsequences - 1; m++)
    {
      if (ms->markpos[m - 1] + 1 == ms->markpos[m])
      {
        return fail("sequence %lu is empty", (unsigned long) (m - 1));
      }
    }
    if (ms->markpos[0] == 0)
    {
      return fail("sequence 0 is empty");
    }
    if (ms->markpos[ms->numofsequences - 2] == out - 1)
    {
      return fail("sequence %lu is empty", (unsigned long) (ms->numofsequences - 1));
    }
  }
  return 0;
}

static int parsefastathreaded(Multiseq *ms, const Alphabet *alpha, const Options *o, const uint8_t *input,
                              uint64_t inputlen, int wantois);

/* multiple FASTA -> tis, ois, descriptions, separator positions, special statistics
   (src/kurtz-basic/multiseq-adv.c:1036-1250 readmultiplefastafilesep, :1260-1291 readmultiplefastafileagain) */
static int parsefasta(Multiseq *ms, const Alphabet *alpha, const Options *o, const uint8_t *input, uint64_t inputlen,
                      uint64_t *concatfilesep, int wantois)
{
  uint64_t out = 0, linenum = 1, desccap = 0, sdcap = 0, sdcnt = 0, mpcap = 0, mpcnt = 0;
  unsigned sepnum = 0;
  int indesc = 0, lastwasspecial = 0;
  uint8_t letter[UCHAR_MAX + 1];  /* code of an ordinary letter, 255 for everything that needs a closer look */
  uint8_t *ois = NULL;
  if (parsefastathreaded(ms, alpha, o, input, inputlen, wantois))
  {
    return finishfasta(ms);
  }
  ms->tis = malloc(inputlen + 1);
  if (ms->tis == NULL)
  {
    return fail("not enough memory for the sequence");
  }
  if (wantois)
  {
    /* the original sequence has an entry wherever the transformed one has: both are written in the same sweep */
    ois = ms->ois = malloc(inputlen + 1);
    if (ois == NULL)
    {
      return fail("not enough memory for the original sequence");
    }
  }
  for (unsigned c = 0; c <= UCHAR_MAX; c++)
  {
    const unsigned acode = alpha->symbolmap[c];
    letter[c] = (c != '>' && !isspace((int) c) && acode != UNDEFCHAR && !ISSPECIAL(acode)) ? (uint8_t) acode : 255;
  }
  for (uint64_t i = 0; i < inputlen; i++)
  {
    const uint8_t current = input[i];
    if (!indesc && letter[current] != 255)
    {
      /* a run of ordinary letters (the bulk of any input) */
      uint8_t *t = ms->tis + out;
      uint64_t j = i;
      if (ois != NULL)
      {
        uint8_t *q = ois + out;
        for (; j < inputlen && letter[input[j]] != 255; j++)
        {
          *t++ = letter[input[j]];
          *q++ = input[j];
        }
      } else
      {
        for (; j < inputlen && letter[input[j]] != 255; j++)
        {
          *t++ = letter[input[j]];
        }
      }
      out += j - i;
      lastwasspecial = 0;
      i = j - 1;
      continue;
    }
    if (indesc)
    {
      PUSH(ms->desc, ms->desclen, desccap, current);
      if (current == '\n')
      {
        linenum++;
        indesc = 0;
      }
    } else if (current == '>')
    {
      if (i > 0 && input[i - 1] != '\n')
      {
        return fail("Illegal character '%c' in file \"%s\" line %lu", current, o->files[sepnum],
                    (unsigned long) linenum);
      }
      PUSH(ms->startdesc, sdcnt, sdcap, ms->desclen);
      if (ms->numofsequences > 0)
      {
        if (sepnum + 1 < o->numoffiles && concatfilesep[sepnum] == i)
        {
          ms->filesep[sepnum++] = out;
          linenum = 0;
        }
        PUSH(ms->markpos, mpcnt, mpcap, out);
        if (!lastwasspecial)
        {
          ms->specialranges++;
        }
        ms->specialcharacters++;
        lastwasspecial = 1;
        if (ois != NULL)
        {
          ois[out] = SEPARATOR;
        }
        ms->tis[out++] = SEPARATOR;
      }
      ms->numofsequences++;
      indesc = 1;
    } else if (isspace(current))
    {
      if (current == '\n')
      {
        linenum++;
      }
    } else
    {
      const unsigned acode = alpha->symbolmap[current];
      if (acode == UNDEFCHAR)
      {
        return fail("Illegal character '%c' in file \"%s\" line %lu", current, o->files[sepnum],
                    (unsigned long) linenum);
      }
      if (ISSPECIAL(acode))
      {
        if (!lastwasspecial)
        {
          ms->specialranges++;
        }
        ms->specialcharacters++;
        lastwasspecial = 1;
      } else
      {
        lastwasspecial = 0;
      }
      if (ois != NULL)
      {
        ois[out] = current;
      }
      ms->tis[out++] = (uint8_t) acode;
    }
  }
  PUSH(ms->startdesc, sdcnt, sdcap, ms->desclen);
  ms->totallength = out;
  return finishfasta(ms);
}

/* ---- the same parse by several threads ------------------------------------------------------
   A '>' at the start of a line is always the start of a record (a description ends at its newline), so the input can
   be cut at such places and every piece parsed on its own: parsefastapiece is parsefasta's state machine without the
   parts that need the whole file (several input files, line numbers for messages).  Each piece writes its symbols into
   the shared tis / ois buffers at the offset of its own input (a piece never produces more symbols than it has bytes),
   the pieces are moved together afterwards, descriptions, separator positions and the special statistics are joined.
   Anything a piece cannot handle (an illegal character, memory) makes the caller fall back to parsefasta, which then
   produces the reference's message.  MKVTREE_PARSE_THREADS (default: usable CPUs, at most 8; 1 = off) and
   MKVTREE_PARSE_MIN (default 4 MB of input) control it; tests/test_cli_fuzz.py runs both parsers against each other and
   against the reference binary. */
typedef struct
{
  const uint8_t *input;
  uint64_t begin, end;            /* this piece of the input; input[begin] == '>' at the start of a line */
  const uint8_t *letter;          /* see parsefasta */
  const Alphabet *alpha;
  uint8_t *tis, *ois;             /* shared buffers; the piece writes from offset `begin` on */
  int first;                      /* the piece that starts the file: no separator in front of its first record */
  uint64_t out;                   /* symbols produced */
  uint8_t *desc;
  uint64_t desclen, desccap;
  uint64_t *startdesc, sdcnt, sdcap;   /* offsets into this piece's desc */
  uint64_t *markpos, mpcnt, mpcap;     /* separator positions, relative to the piece's first symbol */
  uint64_t numofsequences, specialcharacters, specialranges;
  int lastwasspecial;             /* state behind the last symbol of the piece */
  int failed;
  pthread_t thread;
} Fastapiece;

#define PIECEPUSH(ARR, CNT, CAP, VAL)                              \
  do                                                               \
  {                                                                \
    if ((CNT) == (CAP))                                            \
    {                                                              \
      void *grown__;                                               \
      (CAP) = (CAP) ? 2 * (CAP) : 1024;                            \
      grown__ = realloc((ARR), (CAP) * sizeof *(ARR));             \
      if (grown__ == NULL)                                         \
      {                                                            \
        pc->failed = 1;                                            \
        return NULL;                                               \
      }                                                            \
      (ARR) = grown__;                                             \
    }                                                              \
    (ARR)[(CNT)++] = (VAL);                                        \
  } while (0)

static void *parsefastapiece(void *arg)
{
  Fastapiece *pc = (Fastapiece *) arg;
  const uint8_t *input = pc->input, *letter = pc->letter;
  uint8_t *tis = pc->tis + pc->begin, *ois = pc->ois != NULL ? pc->ois + pc->begin : NULL;
  uint64_t out = 0;
  int indesc = 0, lastwasspecial = 0;

  for (uint64_t i = pc->begin; i < pc->end; i++)
  {
    const uint8_t current = input[i];
    if (!indesc && letter[current] != 255)
    {
      uint64_t j = i;
      if (ois != NULL)
      {
        for (; j < pc->end && letter[input[j]] != 255; j++)
        {
          tis[out + (j - i)] = letter[input[j]];
          ois[out + (j - i)] = input[j];
        }
      } else
      {
        for (; j < pc->end && letter[input[j]] != 255; j++)
        {
          tis[out + (j - i)] = letter[input[j]];
        }
      }
      out += j - i;
      lastwasspecial = 0;
      i = j - 1;
      continue;
    }
    if (indesc)
    {
      PIECEPUSH(pc->desc, pc->desclen, pc->desccap, current);
      if (current == '\n')
      {
        indesc = 0;
      }
    } else if (current == '>')
    {
      if (i > 0 && input[i - 1] != '\n')
      {
        pc->failed = 1;  /* the sequential parser reports it */
        return NULL;
      }
      PIECEPUSH(pc->startdesc, pc->sdcnt, pc->sdcap, pc->desclen);
      if (!pc->first || pc->numofsequences > 0)
      {
        PIECEPUSH(pc->markpos, pc->mpcnt, pc->mpcap, out);
        if (!lastwasspecial)
        {
          pc->specialranges++;
        }
        pc->specialcharacters++;
        lastwasspecial = 1;
        if (ois != NULL)
        {
          ois[out] = SEPARATOR;
        }
        tis[out++] = SEPARATOR;
      }
      pc->numofsequences++;
      indesc = 1;
    } else if (!isspace(current))
    {
      const unsigned acode = pc->alpha->symbolmap[current];
      if (acode == UNDEFCHAR)
      {
        pc->failed = 1;
        return NULL;
      }
      if (ISSPECIAL(acode))
      {
        if (!lastwasspecial)
        {
          pc->specialranges++;
        }
        pc->specialcharacters++;
        lastwasspecial = 1;
      } else
      {
        lastwasspecial = 0;
      }
      if (ois != NULL)
      {
        ois[out] = current;
      }
      tis[out++] = (uint8_t) acode;
    }
  }
  pc->out = out;
  pc->lastwasspecial = lastwasspecial;
  return NULL;
}

static unsigned parsethreads(uint64_t inputlen)
{
  const char *e = getenv("MKVTREE_PARSE_THREADS"), *m = getenv("MKVTREE_PARSE_MIN");
  const uint64_t minbytes = m != NULL ? strtoull(m, NULL, 10) : ((uint64_t) 4 << 20);
  long n = e != NULL ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
  if (e == NULL)
  {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0 && CPU_COUNT(&set) < n)
    {
      n = CPU_COUNT(&set);
    }
    if (n > 8)
    {
      n = 8;
    }
  }
  if (n > 64)
  {
    n = 64;
  }
  return (n < 2 || inputlen < minbytes) ? 1u : (unsigned) n;
}

/* 1: parsed by several threads, ms is complete up to totallength (the caller finishes like parsefasta does);
   0: not applicable or a piece gave up, ms untouched, use parsefasta */
static int parsefastathreaded(Multiseq *ms, const Alphabet *alpha, const Options *o, const uint8_t *input,
                              uint64_t inputlen, int wantois)
{
  Fastapiece pcs[64];
  uint8_t letter[UCHAR_MAX + 1];
  unsigned want = parsethreads(inputlen), npieces = 0;
  uint8_t *tis = NULL, *ois = NULL;
  uint64_t prev = 0, outtotal = 0, desctotal = 0, sdtotal = 0, mptotal = 0;
  int ok = 1;

  if (want < 2 || o->numoffiles != 1 || inputlen == 0 || input[0] != '>')
  {
    return 0;
  }
  memset(pcs, 0, sizeof pcs);
  /* pieces start at the first record start at or behind k / want of the input */
  for (unsigned k = 0; k < want; k++)
  {
    uint64_t b = k == 0 ? 0 : inputlen / want * k;
    if (k > 0)
    {
      const uint8_t *hit = NULL;
      if (b <= prev)
      {
        b = prev + 1;
      }
      for (uint64_t p = b; p + 1 < inputlen; p++)
      {
        const uint8_t *nl = memchr(input + p, '\n', inputlen - 1 - p);
        if (nl == NULL)
        {
          break;
        }
        if (nl[1] == '>')
        {
          hit = nl + 1;
          break;
        }
        p = (uint64_t) (nl - input);
      }
      if (hit == NULL)
      {
        break;
      }
      b = (uint64_t) (hit - input);
      pcs[npieces - 1].end = b;
    }
    pcs[npieces].begin = b;
    pcs[npieces].end = inputlen;
    pcs[npieces].first = k == 0;
    prev = b;
    npieces++;
  }
  if (npieces < 2)
  {
    return 0;
  }
  tis = malloc(inputlen + 1);
  ois = wantois ? malloc(inputlen + 1) : NULL;
  if (tis == NULL || (wantois && ois == NULL))
  {
    free(tis);
    free(ois);
    return 0;
  }
  for (unsigned c = 0; c <= UCHAR_MAX; c++)
  {
    const unsigned acode = alpha->symbolmap[c];
    letter[c] = (c != '>' && !isspace((int) c) && acode != UNDEFCHAR && !ISSPECIAL(acode)) ? (uint8_t) acode : 255;
  }
  for (unsigned k = 0; k < npieces; k++)
  {
    pcs[k].input = input;
    pcs[k].letter = letter;
    pcs[k].alpha = alpha;
    pcs[k].tis = tis;
    pcs[k].ois = ois;
  }
  for (unsigned k = 1; k < npieces; k++)
  {
    if (pthread_create(&pcs[k].thread, NULL, parsefastapiece, &pcs[k]) != 0)
    {
      pcs[k].failed = 2;  /* not started */
      ok = 0;
    }
  }
  parsefastapiece(&pcs[0]);
  for (unsigned k = 1; k < npieces; k++)
  {
    if (pcs[k].failed != 2)
    {
      pthread_join(pcs[k].thread, NULL);
    }
  }
  for (unsigned k = 0; k < npieces; k++)
  {
    ok = ok && !pcs[k].failed;
    desctotal += pcs[k].desclen;
    sdtotal += pcs[k].sdcnt;
    mptotal += pcs[k].mpcnt;
  }
  if (ok)
  {
    ms->desc = malloc(desctotal + 1);
    ms->startdesc = malloc((sdtotal + 1) * sizeof *ms->startdesc);
    ms->markpos = malloc((mptotal + 1) * sizeof *ms->markpos);
    ok = ms->desc != NULL && ms->startdesc != NULL && ms->markpos != NULL;
  }
  if (!ok)
  {
    for (unsigned k = 0; k < npieces; k++)
    {
      free(pcs[k].desc);
      free(pcs[k].startdesc);
      free(pcs[k].markpos);
    }
    free(ms->desc);
    free(ms->startdesc);
    free(ms->markpos);
    ms->desc = NULL;
    ms->startdesc = ms->markpos = NULL;
    free(tis);
    free(ois);
    return 0;
  }
  /* join: symbols move left to where the pieces in front of them end (a piece starts at or behind that place) */
  {
    uint64_t sd = 0, mp = 0;
    int lastwasspecial = 0;
    desctotal = 0;
    for (unsigned k = 0; k < npieces; k++)
    {
      Fastapiece *pc = &pcs[k];
      if (k > 0)
      {
        memmove(tis + outtotal, tis + pc->begin, pc->out);
        if (ois != NULL)
        {
          memmove(ois + outtotal, ois + pc->begin, pc->out);
        }
        if (lastwasspecial)
        {
          pc->specialranges--;  /* its leading separator continues the special range the piece in front ended with */
        }
      }
      if (pc->out > 0)
      {
        lastwasspecial = pc->lastwasspecial;
      }
      memcpy(ms->desc + desctotal, pc->desc, pc->desclen);
      for (uint64_t t = 0; t < pc->sdcnt; t++)
      {
        ms->startdesc[sd++] = desctotal + pc->startdesc[t];
      }
      for (uint64_t t = 0; t < pc->mpcnt; t++)
      {
        ms->markpos[mp++] = outtotal + pc->markpos[t];
      }
      ms->numofsequences += pc->numofsequences;
      ms->specialcharacters += pc->specialcharacters;
      ms->specialranges += pc->specialranges;
      desctotal += pc->desclen;
      outtotal += pc->out;
      free(pc->desc);
      free(pc->startdesc);
      free(pc->markpos);
    }
    ms->startdesc[sd] = desctotal;
    ms->desclen = desctotal;
  }
  ms->tis = tis;
  ms->ois = ois;
  ms->totallength = outtotal;
  if (getenv("MKVTREE_PARSE_REPORT") != NULL)
  {
    fprintf(stderr, "parsefasta: %u pieces on %u threads\n", npieces, npieces);
  }
  return 1;
}

/* a file that is not FASTA is taken verbatim as one sequence (src/Mkvtree/mkvinput.c:78-128) */
static int parseplain(Multiseq *ms, const Alphabet *alpha, const uint8_t *input, uint64_t inputlen, int wantois)
{
  static const char dummy[] = "sequence 1\n";
  ms->tis = malloc(inputlen + 1);
  ms->desc = malloc(sizeof dummy);
  ms->startdesc = malloc(2 * sizeof *ms->startdesc);
  if (ms->tis == NULL || ms->desc == NULL || ms->startdesc == NULL)
  {
    return fail("not enough memory for the sequence");
  }
  memcpy(ms->desc, dummy, sizeof dummy - 1);
  ms->desclen = sizeof dummy - 1;
  ms->startdesc[0] = 0;
  ms->startdesc[1] = ms->desclen;
  ms->numofsequences = 1;
  ms->totallength = inputlen;
  if (wantois)
  {
    uint64_t q = 0;
    ms->ois = malloc(inputlen + 1);
    if (ms->ois == NULL)
    {
      return fail("not enough memory for the original sequence");
    }
    for (uint64_t i = 0; i < inputlen; i++)
    {
      if (!isspace(input[i]))
      {
        ms->ois[q++] = input[i];
      }
    }
  }
  for (uint64_t i = 0; i < inputlen; i++)
  {
    const unsigned acode = alpha->symbolmap[input[i]];
    if (acode == UNDEFCHAR)
    {
      switch (input[i])
      {
        case '\n': return fail("Illegal newline");
        case '\t': return fail("Illegal tabulator");
        case ' ': return fail("Illegal blank");
        default:
          return isprint(input[i]) ? fail("Illegal character '%c'", input[i])
                                   : fail("Illegal character '\\%lu'", (unsigned long) input[i]);
      }
    }
    ms->tis[i] = (uint8_t) acode;
  }
  return 0;  /* the reference leaves the special statistics of plain input at zero */
}

/* Plain text without an alphabet option: every byte value is a letter, 254 and 255 included (the reference sorts such
   input with its sorter for sequences without special symbols).  mapped: the bytes are replaced by their rank among
   the bytes that occur.  Otherwise the table holds the bytes themselves and ms->sortseq the ranked copy for the
   sorter (the order of the suffixes is the same). */
static int parsetext(Multiseq *ms, Alphabet *alpha, const uint8_t *input, uint64_t inputlen, int mapped)
{
  int occurs[UCHAR_MAX + 1] = {0};
  uint8_t rank[UCHAR_MAX + 1];
  ms->tis = malloc(inputlen + 1);
  if (ms->tis == NULL)
  {
    return fail("not enough memory for the sequence");
  }
  ms->numofsequences = 1;
  ms->totallength = inputlen;
  for (uint64_t i = 0; i < inputlen; i++)
  {
    occurs[input[i]] = 1;
  }
  alpha->mapsize = 0;
  for (unsigned c = 0; c <= UCHAR_MAX; c++)
  {
    rank[c] = 0;
    if (occurs[c])
    {
      alpha->characters[alpha->mapsize] = (unsigned char) c;
      rank[c] = (uint8_t) alpha->mapsize++;
    }
  }
  alpha->domainsize = alpha->mapsize;
  if (mapped)
  {
    for (uint64_t i = 0; i < inputlen; i++)
    {
      ms->tis[i] = rank[input[i]];
    }
    return 0;
  }
  memcpy(ms->tis, input, inputlen);
  ms->sortseq = malloc(inputlen + 1);
  if (ms->sortseq == NULL)
  {
    return fail("not enough memory for the sequence");
  }
  for (uint64_t i = 0; i < inputlen; i++)
  {
    ms->sortseq[i] = rank[input[i]];
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* output */

/* ------------------------------------------------------------------------------------------------------ */
/* GenBank, EMBL and SwissProt flat files (src/kurtz-basic/parsemultiform.c:225-385): every record becomes a FASTA
   record ">ID DEFINITION-line\nSEQUENCE\n", and the FASTA reader takes it from there.
   A record starts with `first` ("LOCUS" / "ID   "), its identifier is the next blank-delimited word, its description
   the rest of the first line that starts with `second` ("\nDEFINITION" / "\nDE   "), its sequence everything behind
   the line of `third` ("\nORIGIN" / "\nSQ   ") up to "//" without blanks, newlines and digits.
   Returns the length of the FASTA text in `out`, 0 when the buffer does not start with `first`, -1 on error. */
static const uint8_t *findbytes(const uint8_t *text, uint64_t textlen, const char *pattern)
{
  const size_t plen = strlen(pattern);
  if (plen == 0 || textlen < plen)
  {
    return NULL;
  }
  return (const uint8_t *) memmem(text, textlen, pattern, plen);
}

static int64_t flatfiletofasta(uint8_t *out, const uint8_t *db, uint64_t dblen, int checktype, const char *first,
                               const char *second, const char *third)
{
  const uint8_t *end = db + dblen, *rec = db;
  const size_t firstlen = strlen(first);
  uint64_t written = 0;
  while (rec < end)
  {
    const uint8_t *p, *q;
    if ((uint64_t) (end - rec) < firstlen || memcmp(rec, first, firstlen) != 0)
    {
      return 0;  /* (also when a later record does not start with the key: the reference gives up on the format) */
    }
    if (checktype)
    {
      /* EMBL ("... BP.") or SwissProt ("... AA."): the first line tells, and must */
      for (p = rec; p < end && *p != '\n'; p++)
      {
      }
      if (p - rec < 3 || p[-1] != '.'
          || !((tolower(p[-3]) == 'b' && tolower(p[-2]) == 'p') || (tolower(p[-3]) == 'a' && tolower(p[-2]) == 'a')))
      {
        fail("cannot detect sequence type:  first line must end with \"aa.\" or \"bp.\" in lower or upper case");
        return -1;
      }
    }
    for (p = rec + firstlen; p < end && isspace(*p); p++)  /* the identifier */
    {
    }
    q = findbytes(rec, (uint64_t) (end - rec), second);
    if (q == NULL)
    {
      fail("missing \"%s\" in database file", second);
      return -1;
    }
    for (q += strlen(second); q < end && *q == ' '; q++)
    {
    }
    out[written++] = '>';
    for (; p < end && !isspace(*p); p++)
    {
      out[written++] = *p;
    }
    out[written++] = ' ';
    for (p = q; p < end && *p != '\n'; p++)
    {
      out[written++] = *p;
    }
    out[written++] = '\n';
    q = findbytes(q, (uint64_t) (end - q), third);
    if (q == NULL)
    {
      fail("missing \"%s\" in GenBank file", third);
      return -1;
    }
    for (q += strlen(third); q < end && *q != '\n'; q++)  /* the rest of the key's line */
    {
    }
    for (; q < end && *q != '/'; q++)
    {
      if (*q != ' ' && *q != '\n' && !isdigit(*q))
      {
        out[written++] = *q;
      }
    }
    out[written++] = '\n';
    if (q + 1 >= end || q[1] != '/')
    {
      fail("input sequence does no contain second '/'");
      return -1;
    }
    for (; q < end && *q != '\n'; q++)
    {
    }
    for (q = q < end ? q + 1 : end; q < end && isspace(*q); q++)
    {
    }
    rec = q;
  }
  return (int64_t) written;
}

static const char *tabsuffixes[] = {"tis", "ois", "des", "ssp", "bwt", "suf", "lcp", "bck", "sti1", "skp", "cld", "cld1",
                                    "iso", "sti", "cfr", "crf", "lsf", "sds", "llv", "al1", "al2", "prj"};

static int removeexistingindex(const char *indexname)
{
  for (size_t i = 0; i < sizeof tabsuffixes / sizeof tabsuffixes[0]; i++)
  {
    char name[PATH_MAX + 16];
    struct stat st;
    snprintf(name, sizeof name, "%s.%s", indexname, tabsuffixes[i]);
    if (stat(name, &st) == 0)
    {
      if (unlink(name) == -1)
      {
        return fail("cannot remove file \"%s\"", name);
      }
      say("remove file \"%s\"", name);
    }
  }
  return 0;
}

static FILE *openthefile(const char *indexname, const char *suffix, int announce)
{
  char name[PATH_MAX + 16];
  struct stat st;
  FILE *fp;
  snprintf(name, sizeof name, "%s.%s", indexname, suffix);
  if (stat(name, &st) == 0)
  {
    fail("cannot create file \"%s\": it already exists", name);
    return NULL;
  }
  fp = fopen(name, "wb");
  if (fp == NULL)
  {
    fail("cannot open file \"%s\": %s", name, strerror(errno));
    return NULL;
  }
  if (announce)
  {
    say("create file \"%s\"", name);
  }
  return fp;
}

static int writetable(const char *indexname, const char *suffix, const void *data, size_t size, uint64_t count,
                      int announce)
{
  FILE *fp = openthefile(indexname, suffix, announce);
  if (fp == NULL)
  {
    return -1;
  }
  if (count > 0 && fwrite(data, size, count, fp) != count)
  {
    fclose(fp);
    return fail("cannot write %lu items of size %lu: errormsg=\"%s\"", (unsigned long) count, (unsigned long) size,
                strerror(errno));
  }
  if (fclose(fp) != 0)
  {
    return fail("cannot close file \"%s.%s\"", indexname, suffix);
  }
  return 0;
}

/* ---- several GPUs build ONE index (include/vstree_b200.h, "sharded builds", transport B) ------------------------
 * MKVTREE_GPUS=G makes the driver start G threads, one per device, which build the G ranges of the suffix array
 * together; the ranges are written to the index files one after the other, so the files are the same as with one
 * GPU.  (MKVTREE_GPUS_ONE_DEVICE=1 puts all ranks on device MKVTREE_GPU: for tests on a single-GPU machine.) */
#define MAXGPUS 16
typedef struct
{
  esa_ctx *ctx;
  const uint8_t *tis;
  uint64_t n;
  uint32_t numofchars, prefixlength, demand;
  esa_result res;
  esa_stats stats;
  int rc;
  char err[512];
} Rankjob;

/* The suffix table is 8 of the 10 output bytes per suffix.  On one GPU it is never copied on the host: the .suf file is
 * created and mapped as soon as the sequence length is known, its pages are touched by a few threads while the main
 * thread writes the host-side tables and waits for the device, and the library downloads the table straight into the
 * mapping (esa_set_result_allocator).  What is left at the place where the reference writes .suf is the -v line. */
typedef struct
{
  char name[PATH_MAX + 16];
  int fd;
  uint8_t *map;
  uint64_t bytes;
  int used;       /* the library took the mapping */
  pthread_t touch[4];
  int ntouch;
} Mappedtable;

typedef struct
{
  uint8_t *begin, *end;
} Touchjob;
static Touchjob touchjobs[4];

static void *touchmain(void *arg)
{
  Touchjob *j = (Touchjob *) arg;
  for (uint8_t *p = j->begin; p < j->end; p += 4096)
  {
    *(volatile uint8_t *) p = 0;
  }
  return NULL;
}

static void mapoutput(Mappedtable *m, const char *indexname, const char *suffix, uint64_t bytes)
{
  memset(m, 0, sizeof *m);
  m->fd = -1;
  snprintf(m->name, sizeof m->name, "%s.%s", indexname, suffix);
  m->fd = open(m->name, O_RDWR | O_CREAT | O_EXCL, 0666);
  if (m->fd < 0)
  {
    return;  /* the ordinary write path will report what is wrong */
  }
  if (ftruncate(m->fd, (off_t) bytes) != 0
      || (m->map = (uint8_t *) mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, m->fd, 0)) == (uint8_t *) MAP_FAILED)
  {
    close(m->fd);
    unlink(m->name);
    m->fd = -1;
    m->map = NULL;
    return;
  }
  m->bytes = bytes;
  for (int t = 0; t < 4; t++)
  {
    const uint64_t part = ((bytes / 4) + 4095) & ~(uint64_t) 4095;
    touchjobs[t].begin = m->map + (uint64_t) t * part < m->map + bytes ? m->map + (uint64_t) t * part : m->map + bytes;
    touchjobs[t].end = m->map + (uint64_t) (t + 1) * part < m->map + bytes ? m->map + (uint64_t) (t + 1) * part : m->map + bytes;
    if (pthread_create(&m->touch[m->ntouch], NULL, touchmain, &touchjobs[t]) == 0)
    {
      m->ntouch++;
    }
  }
}

static void jointouchers(Mappedtable *m)
{
  for (int t = 0; t < m->ntouch; t++)
  {
    pthread_join(m->touch[t], NULL);
  }
  m->ntouch = 0;
}

static void dropmapping(Mappedtable *m, int keepfile)
{
  jointouchers(m);
  if (m->map != NULL)
  {
    munmap(m->map, m->bytes);
    m->map = NULL;
  }
  if (m->fd >= 0)
  {
    close(m->fd);
    m->fd = -1;
    if (!keepfile)
    {
      unlink(m->name);
    }
  }
}

static void *mappedalloc(void *user, uint32_t table, uint64_t bytes)
{
  Mappedtable *m = (Mappedtable *) user;
  if (table == ESA_SUFTAB && m->map != NULL && bytes == m->bytes)
  {
    m->used = 1;
    return m->map;
  }
  return NULL;
}

/* The CUDA context of a one-shot process costs more than parsing 100 Mbp: esa_init (context, streams, kernel images) runs
 * on its own thread while the main thread reads and transforms the input; joined right before the build.  Nothing is
 * allocated there (pinning gigabytes next to the page-cache traffic of the parser was tried in round 1 and proved
 * erratic).  MKVTREE_TIMING=1 prints the phases of the process on stderr. */
typedef struct
{
  pthread_t thread;
  int started, device, rc;
  esa_ctx *ctx;
  char err[512];
} Earlyinit;

static void *earlyinitmain(void *arg)
{
  Earlyinit *e = (Earlyinit *) arg;
  e->rc = esa_init(&e->ctx, e->device, e->err, sizeof e->err);
  return NULL;
}

static double nowseconds(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static void *rankmain(void *arg)
{
  Rankjob *job = (Rankjob *) arg;
  job->rc = esa_build(job->ctx, job->tis, job->n, job->numofchars, job->prefixlength, job->demand, &job->res,
                      &job->stats, job->err, sizeof job->err);
  return NULL;
}

/* one table, all ranges in rank order; `which` picks the table pointer, counts are per range */
static int writeranges(const char *indexname, const char *suffix, FILE *fp, Rankjob *jobs, int njobs, size_t size,
                       int which)
{
  int r;
  if (fp == NULL && (fp = openthefile(indexname, suffix, 1)) == NULL)
  {
    return -1;
  }
  for (r = 0; r < njobs; r++)
  {
    const esa_result *res = &jobs[r].res;
    const void *data = which == 0 ? (const void *) res->lcptab : which == 1 ? (const void *) res->llv
                     : which == 2 ? (const void *) res->suftab : which == 3 ? (const void *) res->bwttab
                     : (const void *) res->stitab1;
    /* suffix-array tables come in ranges of the suffix array, .sti1 in slices of the text; both in rank order */
    const uint64_t count = which == 1 ? res->nllv : which == 4 ? res->text_count : res->shard_count;
    if (count > 0 && fwrite(data, size, count, fp) != count)
    {
      fclose(fp);
      return fail("cannot write %lu items of size %lu: errormsg=\"%s\"", (unsigned long) count, (unsigned long) size,
                  strerror(errno));
    }
  }
  if (fclose(fp) != 0)
  {
    return fail("cannot close file \"%s.%s\"", indexname, suffix);
  }
  return 0;
}

int main(int argc, char **argv)
{
  Rankjob jobs[MAXGPUS];
  esa_group *group = NULL;
  int njobs = 1;
  Options opt;
  Alphabet alpha;
  Multiseq ms;
  esa_ctx *ctx = NULL;
  esa_result res;
  esa_stats stats;
  char err[1024];
  uint8_t *raw;
  uint64_t rawlen, concatfilesep[MAXNUMBEROFFILES], skip;
  uint32_t numofchars, k, builddemand;
  int rc, dosort, somethingout;
  FILE *lcpfp = NULL;
  Earlyinit early;
  Mappedtable sufmap;
  const int timing = getenv("MKVTREE_TIMING") != NULL;
  const double t_start = nowseconds();
  double t_parsed = 0, t_hosttables = 0, t_built = 0, t_context = 0;

  memset(&early, 0, sizeof early);
  memset(&sufmap, 0, sizeof sufmap);
  sufmap.fd = -1;
  /* a one-shot process: kernel images are loaded with the context (on the thread that starts it behind the parser, not
     one by one at the first launches of the build: 0.45 s of a cold 100 Mbp build), and the tables come back into
     ordinary memory (pinning 1.3 GB costs more than the staged copy); an existing setting of the user wins */
  setenv("CUDA_MODULE_LOADING", "EAGER", 0);
  setenv("ESA_PAGEABLE_RESULTS", "1", 0);
  rc = parseoptions(&opt, argc, argv);
  if (rc == 1)
  {
    return EXIT_SUCCESS;
  }
  if (rc == -1)
  {
    fprintf(stderr, "%s\n", message);
    return EXIT_FAILURE;
  }
  if (rc < 0)
  {
    goto failure;
  }
  if (make_alphabet(&alpha, &opt) != 0)
  {
    goto failure;
  }
  memset(&ms, 0, sizeof ms);
  memset(concatfilesep, 0, sizeof concatfilesep);
  if ((opt.demand & (ESA_BWTTAB | ESA_SUFTAB | ESA_LCPTAB | ESA_BCKTAB | ESA_STI1TAB | ESA_SKPTAB))
      && getenv("MKVTREE_GPUS") == NULL)
  {
    /* a table of the sorter is demanded: the device is needed, start it now (one GPU; several: see below) */
    const char *dev = getenv("MKVTREE_GPU");
    early.device = dev != NULL ? atoi(dev) : 0;
    early.started = pthread_create(&early.thread, NULL, earlyinitmain, &early) == 0;
  }
  if (alpha.fromtext && opt.numoffiles > 1)
  {
    fail("multiple input files require one of the options -smap, -dna, or -protein");
    goto failure;
  }
  if (opt.numoffiles == 1)
  {
    raw = readwholefile(opt.files[0], &rawlen);
    if (raw == NULL)
    {
      goto failure;
    }
    ms.filelength[0] = rawlen;
    say("reading%sfile \"%s\"", hasgzsuffix(opt.files[0]) ? " gzipped " : " ", opt.files[0]);
  } else
  {
    raw = concatfiles(&opt, &ms, concatfilesep, &rawlen);
    if (raw == NULL)
    {
      goto failure;
    }
  }
  for (skip = 0; skip < rawlen && isspace(raw[skip]); skip++)
  {
  }
  if (alpha.fromtext)
  {
    /* no alphabet option: plain text only; with a prefix length its bytes are numbered in byte order (.al2), without
       one they are sorted as they are (src/Mkvtree/mkvinput.c:100-170) */
    if ((skip < rawlen && raw[skip] == '>') || (rawlen - skip >= 5 && memcmp(raw + skip, "LOCUS", 5) == 0)
        || (rawlen >= 5 && memcmp(raw, "ID   ", 5) == 0))
    {
      fail("fasta, EMBL, and GenBank format requires one of the options -smap, -dna, or -protein");
      goto failure;
    }
    if (opt.prefixlength == AUTOPREFIXLENGTH)
    {
      fail("option -pl requires positive integer as parameter since alphabet size is unknown");
      goto failure;
    }
    if (parsetext(&ms, &alpha, raw, rawlen, opt.prefixlength > 0) != 0)
    {
      goto failure;
    }
  } else if (skip < rawlen && raw[skip] == '>')
  {
    /* the separator offsets of concatenated files refer to the unskipped buffer, but a concatenation never
       starts with white space */
    if (parsefasta(&ms, &alpha, &opt, raw + skip, rawlen - skip, concatfilesep, (opt.demand & ESA_OISTAB) != 0) != 0)
    {
      goto failure;
    }
  } else
  {
    /* A flat file has to start with its key: behind leading white space the reference recognises neither format
       properly (EMBL is looked for at the very start of the file, a GenBank conversion is read from the very start,
       src/kurtz-basic/parsemultiform.c:352-385, mkvinput.c:232-256) and ends up rejecting the white space like the
       plain reader below does.  Several input files are FASTA or plain. */
    uint8_t *fasta = NULL;
    int64_t flen = 0;
    if (opt.numoffiles == 1 && rawlen > 0 && skip == 0)
    {
      fasta = malloc(rawlen + 16);
      if (fasta == NULL)
      {
        fail("not enough memory for %lu bytes", (unsigned long) rawlen);
        goto failure;
      }
      flen = flatfiletofasta(fasta, raw + skip, rawlen - skip, 0, "LOCUS", "\nDEFINITION", "\nORIGIN");
      if (flen == 0)
      {
        flen = flatfiletofasta(fasta, raw, rawlen, 1, "ID   ", "\nDE   ", "\nSQ   ");
      }
      if (flen < 0)
      {
        goto failure;
      }
    }
    if (flen > 0)
    {
      if (parsefasta(&ms, &alpha, &opt, fasta, (uint64_t) flen, concatfilesep, (opt.demand & ESA_OISTAB) != 0) != 0)
      {
        goto failure;
      }
    } else if (parseplain(&ms, &alpha, raw, rawlen, (opt.demand & ESA_OISTAB) != 0) != 0)
    {
      goto failure;
    }
    free(fasta);
  }
  free(raw);
  if (ms.totallength == 0)
  {
    fail("input sequences are all empty");
    goto failure;
  }
  if (transformsequence(ms.tis, ms.totallength, opt.isset[OPTREV], opt.isset[OPTCPL]) != 0)
  {
    goto failure;
  }
  numofchars = alpha.fromtext ? alpha.mapsize : alpha.mapsize - 1;
  k = opt.prefixlength;
  if (k == AUTOPREFIXLENGTH)
  {
    k = esa_recommended_prefixlength(numofchars, ms.totallength);
  }
  if (k > 0)
  {
    const uint32_t maxk = esa_maximal_prefixlength(numofchars, ms.totallength);
    if (maxk < k)
    {
      fail("prefix length %lu is too large, maximal prefix length for this input size and alphabet size is %lu",
           (unsigned long) k, (unsigned long) maxk);
      goto failure;
    }
    say("for this input size and alphabet size, the maximal prefixlength\n(argument of option -pl) is %lu,\n"
        "the recommended prefixlength is %lu", (unsigned long) maxk,
        (unsigned long) esa_recommended_prefixlength(numofchars, ms.totallength));
  }
  if (ms.numofsequences > 1)
  {
    say("total length of sequences: %lu (including %lu separator%s)", (unsigned long) ms.totallength,
        (unsigned long) (ms.numofsequences - 1), ms.numofsequences == 2 ? "" : "s");
  } else
  {
    say("total length of sequences: %lu", (unsigned long) ms.totallength);
  }
  if (k > 0)
  {
    verbosealphabet(&alpha);
  }
  t_parsed = nowseconds();
  somethingout = (opt.demand & OUTPUTTABS) != 0;
  if (somethingout && removeexistingindex(opt.indexname) != 0)
  {
    goto failure;
  }
  if ((opt.demand & ESA_SUFTAB) && getenv("MKVTREE_GPUS") == NULL && getenv("MKVTREE_NO_MAPPED_SUF") == NULL)
  {
    mapoutput(&sufmap, opt.indexname, "suf", 8 * ((uint64_t) ms.totallength + 1));
  }
  if (somethingout && ms.numofsequences > 1
      && writetable(opt.indexname, "ssp", ms.markpos, sizeof(uint64_t), ms.numofsequences - 1, 0) != 0)
  {
    goto failure;
  }
  if ((opt.demand & ESA_TISTAB) && writetable(opt.indexname, "tis", ms.tis, 1, ms.totallength, 1) != 0)
  {
    goto failure;
  }
  if ((opt.demand & ESA_OISTAB) && ms.ois != NULL
      && writetable(opt.indexname, "ois", ms.ois, 1, ms.totallength, 1) != 0)
  {
    goto failure;
  }
  if (somethingout && ms.desc != NULL)
  {
    if (writetable(opt.indexname, "des", ms.desc, 1, ms.desclen, 1) != 0
        || writetable(opt.indexname, "sds", ms.startdesc, sizeof(uint64_t), ms.numofsequences + 1, 1) != 0)
    {
      goto failure;
    }
  }
  dosort = !((opt.demand & (ESA_TISTAB | ESA_OISTAB))
             && !(opt.demand & (ESA_BWTTAB | ESA_SUFTAB | ESA_LCPTAB | ESA_BCKTAB | ESA_STI1TAB | ESA_SKPTAB)));
  memset(&res, 0, sizeof res);
  memset(&stats, 0, sizeof stats);
  t_hosttables = nowseconds();
  if (dosort)
  {
    if (opt.demand & ESA_LCPTAB)
    {
      lcpfp = openthefile(opt.indexname, "lcp", 1);  /* the reference opens .lcp before it sorts */
      if (lcpfp == NULL)
      {
        goto failure;
      }
    }
    say("initializing data structures");
    if (k > 0)
    {
      say("sorting suffixes according to prefix of length %lu", (unsigned long) k);
      say("sorting all buckets");
    } else
    {
      say("sorting suffixes");
    }
    builddemand = ESA_SUFTAB | (opt.demand & (ESA_BWTTAB | ESA_LCPTAB | ESA_BCKTAB | ESA_STI1TAB | ESA_SKPTAB));
    {
      const char *dev = getenv("MKVTREE_GPU"), *many = getenv("MKVTREE_GPUS"), *one = getenv("MKVTREE_GPUS_ONE_DEVICE");
      const int first = dev != NULL ? atoi(dev) : 0;
      int r;
      njobs = many != NULL ? atoi(many) : 1;
      if (njobs < 1 || njobs > MAXGPUS)
      {
        fail("MKVTREE_GPUS must be in the range 1..%d", MAXGPUS);
        goto failure;
      }
      memset(jobs, 0, sizeof jobs);
      if (njobs > 1 && esa_group_create(&group, njobs, err, sizeof err) != 0)
      {
        fail("%s", err);
        goto failure;
      }
      for (r = 0; r < njobs; r++)
      {
        const int device = (one != NULL && atoi(one) != 0) ? first : first + r;
        if (r == 0 && early.started)
        {
          pthread_join(early.thread, NULL);  /* the context started behind the parser */
          early.started = 0;
          if (early.rc != 0)
          {
            fail("%s", early.err);
            goto failure;
          }
          jobs[0].ctx = early.ctx;
          early.ctx = NULL;
          t_context = nowseconds();
        } else if (esa_init(&jobs[r].ctx, device, err, sizeof err) != 0)
        {
          fail("%s", err);
          goto failure;
        }
        if (njobs > 1 && esa_comm_attach_group(jobs[r].ctx, group, r, err, sizeof err) != 0)
        {
          fail("%s", err);
          goto failure;
        }
        jobs[r].tis = ms.sortseq != NULL ? ms.sortseq : ms.tis;
        jobs[r].n = ms.totallength;
        jobs[r].numofchars = numofchars;
        jobs[r].prefixlength = k;
        jobs[r].demand = builddemand;
      }
      if (njobs == 1)
      {
        if (sufmap.map != NULL)
        {
          jointouchers(&sufmap);
          esa_set_result_allocator(jobs[0].ctx, mappedalloc, &sufmap);
        }
        rankmain(&jobs[0]);
      } else
      {
        pthread_t threads[MAXGPUS];
        for (r = 0; r < njobs; r++)
        {
          if (pthread_create(&threads[r], NULL, rankmain, &jobs[r]) != 0)
          {
            fail("cannot start the thread of GPU %d", r);
            esa_group_abort(group);
            njobs = r;
            break;
          }
        }
        for (r = 0; r < njobs; r++)
        {
          pthread_join(threads[r], NULL);
        }
      }
      for (r = 0; r < njobs; r++)
      {
        if (jobs[r].rc != 0 && jobs[r].rc != ESA_ERR_COMM)
        {
          fail("%s", jobs[r].err);  /* the rank that failed first; the others only report that a peer failed */
          goto failure;
        }
      }
      for (r = 0; r < njobs; r++)
      {
        if (jobs[r].rc != 0)
        {
          fail("%s", jobs[r].err);
          goto failure;
        }
      }
      if (ms.sortseq != NULL)
      {
        /* raw text without an alphabet: the sorter saw ranked bytes, the Burrows-Wheeler table holds text bytes */
        for (r = 0; r < njobs; r++)
        {
          uint8_t *bwt = jobs[r].res.bwttab;
          for (uint64_t i = 0; bwt != NULL && i < jobs[r].res.shard_count; i++)
          {
            if (bwt[i] < ESA_UNDEFBWTCHAR)
            {
              bwt[i] = alpha.characters[bwt[i]];
            }
          }
        }
      }
      ctx = jobs[0].ctx;
      res = jobs[0].res;       /* bcktab, longest, maxbranchdepth are complete on every rank */
      stats = jobs[0].stats;
      for (r = 1; r < njobs; r++)
      {
        res.nllv += jobs[r].res.nllv;
        stats.d2h_bytes += jobs[r].stats.d2h_bytes;
        if (jobs[r].stats.device_bytes_peak > stats.device_bytes_peak)
        {
          stats.device_bytes_peak = jobs[r].stats.device_bytes_peak;
        }
      }
    }
    t_built = nowseconds();
    if (lcpfp != NULL)
    {
      FILE *fp = lcpfp;
      lcpfp = NULL;
      if (writeranges(opt.indexname, "lcp", fp, jobs, njobs, 1, 0) != 0
          || writeranges(opt.indexname, "llv", NULL, jobs, njobs, sizeof(esa_pairuint), 1) != 0)
      {
        goto failure;
      }
    }
    if ((opt.demand & ESA_BCKTAB) && k > 0
        && writetable(opt.indexname, "bck", res.bcktab, sizeof(uint64_t), 2 * res.numofcodes, 1) != 0)
    {
      goto failure;
    }
    if ((opt.demand & ESA_SUFTAB) && sufmap.used)
    {
      say("create file \"%s\"", sufmap.name);  /* the table is in the mapped file already */
      dropmapping(&sufmap, 1);
      jobs[0].res.suftab = NULL;
    } else
    {
      dropmapping(&sufmap, 0);
      if ((opt.demand & ESA_SUFTAB) && writeranges(opt.indexname, "suf", NULL, jobs, njobs, sizeof(uint64_t), 2) != 0)
      {
        goto failure;
      }
    }
    if ((opt.demand & ESA_STI1TAB) && writeranges(opt.indexname, "sti1", NULL, jobs, njobs, 1, 4) != 0)
    {
      goto failure;
    }
    if ((opt.demand & ESA_BWTTAB) && writeranges(opt.indexname, "bwt", NULL, jobs, njobs, 1, 3) != 0)
    {
      goto failure;
    }
  }
  if (somethingout)
  {
    uint64_t startseq = 0;
    FILE *fp = openthefile(opt.indexname, "prj", 1);
    if (fp == NULL)
    {
      goto failure;
    }
    for (unsigned f = 0; f < opt.numoffiles; f++)
    {
      fprintf(fp, "%s=%s %lu ", f < opt.numoffiles - opt.numofqueryfiles ? "dbfile" : "queryfile", opt.files[f],
              (unsigned long) ms.filelength[f]);
      if (f == opt.numoffiles - 1)
      {
        fprintf(fp, "%lu\n", (unsigned long) (ms.totallength - startseq));
      } else
      {
        fprintf(fp, "%lu\n", (unsigned long) (ms.filesep[f] - startseq));
        startseq = ms.filesep[f] + 1;
      }
    }
    fprintf(fp, "totallength=%lu\n", (unsigned long) ms.totallength);
    fprintf(fp, "specialcharacters=%lu\n", (unsigned long) ms.specialcharacters);
    fprintf(fp, "specialranges=%lu\n", (unsigned long) ms.specialranges);
    fprintf(fp, "lengthofspecialprefix=%lu\n", (unsigned long) ms.lengthofspecialprefix);
    fprintf(fp, "lengthofspecialsuffix=%lu\n", (unsigned long) ms.lengthofspecialsuffix);
    fprintf(fp, "numofsequences=%lu\n", (unsigned long) ms.numofsequences);
    {
      /* the sequences in front of the separator that ends the last database file are database sequences */
      uint64_t numofdbsequences = ms.numofsequences;
      if (opt.numofqueryfiles > 0)
      {
        const uint64_t lastdbsep = ms.filesep[opt.numoffiles - opt.numofqueryfiles - 1];
        numofdbsequences = 0;
        while (numofdbsequences + 1 < ms.numofsequences && ms.markpos[numofdbsequences] < lastdbsep)
        {
          numofdbsequences++;
        }
        numofdbsequences++;
      }
      fprintf(fp, "numofdbsequences=%lu\n", (unsigned long) numofdbsequences);
      fprintf(fp, "numofquerysequences=%lu\n", (unsigned long) (ms.numofsequences - numofdbsequences));
    }
    if (dosort)
    {
      fprintf(fp, "longest=%lu\n", (unsigned long) res.longest);
    }
    fprintf(fp, "prefixlength=%lu\n", (unsigned long) k);
    if (opt.demand & ESA_LCPTAB)
    {
      fprintf(fp, "largelcpvalues=%lu\n", (unsigned long) res.nllv);
      fprintf(fp, "maxbranchdepth=%lu\n", (unsigned long) res.maxbranchdepth);
    } else
    {
      fprintf(fp, "largelcpvalues=0\nmaxbranchdepth=0\n");
    }
    fprintf(fp, "integersize=64\nlittleendian=1\n");
    if (fclose(fp) != 0)
    {
      fail("cannot close file \"%s.prj\"", opt.indexname);
      goto failure;
    }
    if (alpha.fromtext)
    {
      /* .al2: the size of the alphabet in one byte, then its characters (src/Mkvtree/mkvprocess.c:554-577); no
         alphabet file at all without a prefix length */
      if (k > 0)
      {
        fp = openthefile(opt.indexname, "al2", 1);
        if (fp == NULL)
        {
          goto failure;
        }
        putc((int) (alpha.mapsize & 0xffu), fp);
        fwrite(alpha.characters, 1, alpha.mapsize, fp);
        fclose(fp);
      }
      goto alphabetwritten;
    }
    fp = openthefile(opt.indexname, "al1", 1);
    if (fp == NULL)
    {
      goto failure;
    }
    if (opt.isset[OPTSMAP])
    {
      FILE *in = scanpathsforfile("MKVTREESMAPDIR", opt.smapfile);
      int cc;
      if (in == NULL)
      {
        goto failure;
      }
      while ((cc = fgetc(in)) != EOF)
      {
        putc(cc, fp);
      }
      fclose(in);
    } else if (opt.isset[OPTDNA])
    {
      fputs("aA\ncC\ngG\ntTuU\nnsywrkvbdhmNSYWRKVBDHM\n", fp);
    } else
    {
      fputs("L\nV\nI\nF\nK\nR\nE\nD\nA\nG\nS\nT\nN\nQ\nY\nW\nP\nH\nM\nC\nXUBZ*-\n", fp);
    }
    fclose(fp);
alphabetwritten:;
  }
  /* makeskptab runs last in the reference (src/Mkvtree/mkvprocess.c:767-815, called at :1075 after the project
     and alphabet files are out), so the -v transcript names .skp last */
  if ((opt.demand & ESA_SKPTAB) && res.skiptab != NULL
      && writetable(opt.indexname, "skp", res.skiptab, sizeof(uint64_t), ms.totallength + 1, 1) != 0)
  {
    goto failure;
  }
  if (verbose)
  {
    /* same line format as src/Mkvtree/mkvfile.c:22-34; the numbers are this build's: host tables, device arena */
    const double hostbytes = (double) ms.totallength * ((opt.demand & ESA_OISTAB) ? 2.0 : 1.0)
                             + (double) stats.d2h_bytes;
    printf("overall space peak: main=%.2f MB (%.2f bytes/symbol), secondary=%.2f MB\n", hostbytes / (1 << 20),
           hostbytes / (double) ms.totallength, (double) stats.device_bytes_peak / (1 << 20));
  }
  {
    const double t_written = nowseconds();
    if (ctx != NULL && getenv("MKVTREE_TEARDOWN") == NULL && !timing)
    {
      /* every file is closed: leave the release of the device context and of a gigabyte of tables to the operating
         system (0.14 s of a 100 Mbp run when done call by call) */
      fflush(NULL);
      _exit(EXIT_SUCCESS);
    }
    if (ctx != NULL)
    {
      int r;
      for (r = 0; r < njobs; r++)
      {
        esa_free(jobs[r].ctx, &jobs[r].res);
        esa_shutdown(jobs[r].ctx);
      }
      esa_group_destroy(group);
    }
    if (timing)
    {
      fprintf(stderr, "MKVTREE_TIMING parse %.3f s, host tables %.3f s, wait for context %.3f s, esa_build %.3f s [upload "
                      "%.3f, build on device %.3f, download tail %.3f], table output %.3f s, teardown %.3f s, total %.3f s\n",
              t_parsed - t_start, t_hosttables - t_parsed, t_context > 0 ? t_context - t_hosttables : 0.0,
              dosort ? t_built - (t_context > 0 ? t_context : t_hosttables) : 0.0, stats.ms_h2d * 1e-3,
              stats.ms_total * 1e-3, stats.ms_d2h * 1e-3, dosort ? t_written - t_built : t_written - t_hosttables,
              nowseconds() - t_written, nowseconds() - t_start);
    }
  }
  return EXIT_SUCCESS;

failure:
  dropmapping(&sufmap, 0);
  fprintf(stderr, "%s: %s\n", argv[0], message);
  return EXIT_FAILURE;
}
