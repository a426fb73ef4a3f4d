/* Hand-written replacement for the header the reference generates with bin/vmrelease.sh
   (consumed by kurtz/programversion.c). Only used by oracle/build_ref.sh. */
#define VSTREECOMPILEDATE "oracle"
#define VSTREECFLAGS "(oracle build)"
#define VSTREERELEASEDATE "2007-08-27"
#define VSTREEVERSION "2.3.1"
#define VSTREECOMPILEHOST "oracle"
