/* TEST INFRASTRUCTURE ONLY - force-included (-include) in front of oracle/ref_harness_stats.cpp so
 * that the reference's source/LatencyHistogram.h compiles without boost: the two heavy headers it
 * pulls in (ProgArgs.h, workers/WorkersSharedData.h) are switched off through their include guards
 * and the one boost name of its declarations is forward declared. The inline parts of the class
 * (addLatency, getters, percentiles, operator+=, the two string formatters) are what the harness
 * uses; the boost::property_tree members in LatencyHistogram.cpp are not linked. */
#pragma once
#include <cstdint>
#include <iomanip>
#include <ios>
#include <sstream>
#include <string>
#include "Common.h" /* the shim next to this file: IF_UNLIKELY / IF_LIKELY */
#define PROGARGS_H_
#define WORKERS_WORKERSSHAREDDATA_H_
namespace bpt { class ptree; }
