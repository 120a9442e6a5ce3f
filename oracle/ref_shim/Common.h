/* Shim so that the reference's offset generator headers compile without ProgArgs/boost
 * (SURVEY.md Appendix C). Only used to build oracle/_ref/libelb_ref.so. */
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>
#define IF_UNLIKELY(condition) if(__builtin_expect(!!(condition), 0) )
#define IF_LIKELY(condition) if(__builtin_expect(!!(condition), 1) )
