/* TEST INFRASTRUCTURE ONLY - stand-in for <boost/config.hpp> (absent in this image) so that the
 * reference's source/Common.h compiles where it lies: it only needs the two branch hint macros. */
#pragma once
#define BOOST_LIKELY(x) __builtin_expect(!!(x), 1)
#define BOOST_UNLIKELY(x) __builtin_expect(!!(x), 0)
