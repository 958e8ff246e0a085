// Host constants shared by the C ABI translation units (defined in tables.cpp).
#pragma once
#include <stdint.h>
#include "../../include/hsearch_b200.h"
namespace hs {
extern const double kCoordinates[HS_AA][HS_CDIM];
extern const int kBlosum62[HS_AA][HS_AA];
extern const char kAA20[21];
extern const char kCodeLetters[21];
extern const int kBase[26];
extern const int kReduced[26];
void set_error(const char *fmt, ...);
const char *get_error();
void coordinates_table(uint32_t variant, double *out160);
void blosum_metric(int32_t *out400);
bool blosum_filter_embedding(double *out160);
}  // namespace hs
