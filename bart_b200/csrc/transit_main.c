/* transit_main.c -- the stand-alone CLI, `transit -c cfg [--justOpacity]`, as BART.py:563-565,
 * 632-634 and code/bestFit.py:421-427 invoke it.  Mirrors main() of the reference
 * (modules/transit/transit/src/transit.c:230-242): init, one model from the atmosphere file's
 * own profiles, free.  The spectrum is written to `outspec` in the reference's two-column text
 * format (eclipse.c:355-380 / slantpath.c:510-555) -- only here, never inside the MCMC loop. */
#include <stdio.h>
#include <stdlib.h>
#include "bart_b200.h"

int bart_cli_run(void);   /* in libbart_b200: runs the atmosphere-file model, writes outspec */

int main(int argc, char **argv) {
  transit_init(argc, argv);
  int rc = bart_cli_run();
  free_memory();
  return rc == 0 ? EXIT_SUCCESS : EXIT_FAILURE;
}
