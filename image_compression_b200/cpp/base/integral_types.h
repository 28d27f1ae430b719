// Fixed-width integer names used throughout the image_compression public API.
// Interface-compatible stand-in for the reference's base/integral_types.h (same names, <stdint.h> underneath).
#ifndef BASE_INTEGRAL_TYPES_H_
#define BASE_INTEGRAL_TYPES_H_

#include <stdint.h>

typedef int8_t int8;
typedef int16_t int16;
typedef int32_t int32;
typedef int64_t int64;
typedef uint8_t uint8;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef uint64_t uint64;

static const uint8 kuint8max = 0xFF;
static const uint16 kuint16max = 0xFFFF;
static const uint32 kuint32max = 0xFFFFFFFFu;
static const uint64 kuint64max = 0xFFFFFFFFFFFFFFFFull;
static const int8 kint8min = -0x7F - 1;
static const int8 kint8max = 0x7F;
static const int16 kint16min = -0x7FFF - 1;
static const int16 kint16max = 0x7FFF;
static const int32 kint32min = -0x7FFFFFFF - 1;
static const int32 kint32max = 0x7FFFFFFF;
static const int64 kint64min = -0x7FFFFFFFFFFFFFFFll - 1;
static const int64 kint64max = 0x7FFFFFFFFFFFFFFFll;

#endif  // BASE_INTEGRAL_TYPES_H_
