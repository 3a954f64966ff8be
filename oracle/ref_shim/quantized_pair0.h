/* quantized_pair0.h -- stand-in for the reference's missing large blob
 * (.MISSING_LARGE_BLOBS:1).  Same identifiers and shapes as the generated header
 * (python/superpoint_inference.py:630-664; include/data/quantized/quantized_image0.h:
 * 5-15,1938-1939), but the arrays are mutable globals that oracle/ref_harness.c
 * fills before each run, so one build of the unmodified tracking_main.c serves any
 * input at the reference's native 24x80 shape.  TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <stdint.h>

extern int cell_size;

extern int image0_rows, image0_cols, image0_channels;
extern int image0_feature_rows, image0_feature_cols;
extern float image0_semi_scale;
extern int8_t image0_semi[1920][65];
extern float image0_desc_scale;
extern int8_t image0_desc[1920][256];

extern int image1_rows, image1_cols, image1_channels;
extern int image1_feature_rows, image1_feature_cols;
extern float image1_semi_scale;
extern int8_t image1_semi[1920][65];
extern float image1_desc_scale;
extern int8_t image1_desc[1920][256];
