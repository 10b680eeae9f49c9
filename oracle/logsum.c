/* logsum.c -- ORACLE (test infrastructure only).  Follows src/logsum.c:58-111. */
#include <math.h>
#include "bath_oracle.h"

#define BO_LOGSUM_SCALE 1000.f
#define BO_LOGSUM_TBL   16000

static float flogsum_lookup[BO_LOGSUM_TBL];
static int   flogsum_ready = 0;

/* src/logsum.c:80-91 */
void bo_FLogsumInit(void)
{
  int i;
  if (flogsum_ready) return;
  for (i = 0; i < BO_LOGSUM_TBL; i++)
    flogsum_lookup[i] = log(1. + exp((double) -i / BO_LOGSUM_SCALE));
  flogsum_ready = 1;
}

/* src/logsum.c:104-111 */
float bo_FLogsum(float a, float b)
{
  const float max = (a > b) ? a : b;
  const float min = (a > b) ? b : a;
  if (!flogsum_ready) bo_FLogsumInit();
  return (min == -INFINITY || (max - min) >= 15.7f) ? max
         : max + flogsum_lookup[(int)((max - min) * BO_LOGSUM_SCALE)];
}
