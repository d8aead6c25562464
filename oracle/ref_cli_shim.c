/*
 * TEST INFRASTRUCTURE - not part of the product.
 * The UNMODIFIED reference do_quantsmooth (scalar build) under its public name, so that
 * oracle/Makefile can link the product's CLI front end (csrc/jpegqs.c + csrc/jpegcoef.c) against
 * the reference instead of the CUDA back end: oracle/_ref/jpegqs_ref is then "the reference
 * tool with the same JPEG codec", and its output files are what the CUDA-backed `jpegqs` must
 * reproduce byte for byte (tests/test_cli.py).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "jpeglib.h"
#define logfmt(...) fprintf(stderr, __VA_ARGS__)
#define WITH_LOG
#define TRANSCODE_ONLY
#define JPEGQS_ATTR
#include "quantsmooth.h"
