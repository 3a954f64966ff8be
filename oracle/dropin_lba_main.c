/* dropin_lba_main.c -- wrapper of the drop-in demonstration for src/local_bundle_adjustment.c
 * (oracle/Makefile, target dropin).  That file is compiled unmodified (main() renamed) against THIS
 * repository's include/gemmini_functions_cpu.h, which only declares matmul/matmul2, so its 8 500
 * matmul2 calls per run resolve to libmaveric_b200.so and execute on the GPU.  Its stub cholesky()
 * is weakened; this one prints the matrix it receives, one hex word per entry.
 * TEST INFRASTRUCTURE ONLY. */
#include <stdio.h>
#include <string.h>
int lba_main(void);
void cholesky(float* matrix, int dim, int stride) {
  (void)stride;
  printf("dim %d\n", dim);
  for (int i = 0; i < dim * dim; i++) {
    unsigned u;
    memcpy(&u, matrix + i, 4);
    printf("%08x\n", u);
  }
}
int main(void) { return lba_main(); }
