/* Same tunables header as the reference (parameters.h:1-2).  FACTOR is printed and
 * never used there; CUTOFF was the OpenMP `if` threshold (omp_smithW.c:209) and has no
 * meaning on the GPU path.  Both are kept so that the CLI's first line and the
 * reference's tuneCutoff.sh-style scripts keep working. */
#define FACTOR 128
#define CUTOFF 1024
