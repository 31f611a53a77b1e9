#pragma once
/* stub (see hdk/stubs/README.md) */
#include <cstdint>
#include <cstddef>
#include <memory>
#include <string>
#include <vector>
typedef int64_t exint;
typedef float fpreal32;
typedef double fpreal64;
struct fpreal16 { uint16_t bits; };           /* IEEE binary16 storage, as in SYS/fpreal16.h */
