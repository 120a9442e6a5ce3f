/*
 * TEST INFRASTRUCTURE ONLY - C ABI around the REFERENCE's own LatencyHistogram.h and UnitTk
 * (toolkits/UnitTk.{h,cpp}), compiled from /root/reference/source where they lie into
 * oracle/_ref/libelb_ref.so (see oracle/Makefile and ref_shim/stats_prelude.h). Pins the oracle's
 * and the product's histogram arithmetic, percentile rule, per-second rule and the human-readable
 * number formats of the result table.
 */
#include <cstring>
#include <string>

#define private public
#include "LatencyHistogram.h"
#undef private
#include "ProgException.h"
#include "toolkits/UnitTk.h"
#include "toolkits/HashTk.h"

static int64_t copyOut(const std::string& text, char* outBuf, uint64_t outBufLen)
{
	if(outBuf && outBufLen)
	{
		const size_t copyLen = (text.size() < (outBufLen - 1) ) ? text.size() : (outBufLen - 1);
		memcpy(outBuf, text.data(), copyLen);
		outBuf[copyLen] = 0;
	}

	return (int64_t)text.size();
}

extern "C" {

void* ref_histogram_create() { return new LatencyHistogram(); }
void ref_histogram_destroy(void* h) { delete (LatencyHistogram*)h; }
void ref_histogram_add(void* h, uint64_t latencyMicroSec)
	{ ( (LatencyHistogram*)h)->addLatency(latencyMicroSec); }
void ref_histogram_merge(void* dst, const void* src)
	{ *(LatencyHistogram*)dst += *(const LatencyHistogram*)src; }
uint64_t ref_histogram_num(const void* h) { return ( (const LatencyHistogram*)h)->getNumStoredValues(); }
uint64_t ref_histogram_min(const void* h) { return ( (const LatencyHistogram*)h)->getMinMicroSecLat(); }
uint64_t ref_histogram_max(const void* h) { return ( (const LatencyHistogram*)h)->getMaxMicroSecLat(); }
uint64_t ref_histogram_avg(const void* h) { return ( (const LatencyHistogram*)h)->getAverageMicroSec(); }
uint64_t ref_histogram_sum(const void* h) { return ( (const LatencyHistogram*)h)->numMicroSecTotal; }
int ref_histogram_exceeded(const void* h) { return ( (const LatencyHistogram*)h)->getHistogramExceeded(); }
double ref_histogram_percentile(const void* h, double percentage)
	{ return ( (const LatencyHistogram*)h)->getPercentile(percentage); }

uint64_t ref_histogram_num_buckets(const void* h)
	{ return ( (const LatencyHistogram*)h)->buckets.size(); }

void ref_histogram_buckets(const void* h, uint64_t* outBuckets)
{
	const LatencyHistogram* histo = (const LatencyHistogram*)h;

	for(size_t i = 0; i < histo->buckets.size(); i++)
		outBuckets[i] = histo->buckets[i];
}

int64_t ref_histogram_str(const void* h, char* outBuf, uint64_t outBufLen)
	{ return copyOut( ( (const LatencyHistogram*)h)->getHistogramStr(), outBuf, outBufLen); }

int64_t ref_histogram_percentile_str(const void* h, double percentage, char* outBuf,
	uint64_t outBufLen)
	{ return copyOut( ( (const LatencyHistogram*)h)->getPercentileStr(percentage), outBuf, outBufLen); }

uint64_t ref_per_sec_from_usec(uint64_t totalValue, uint64_t elapsedUSec)
	{ return UnitTk::getPerSecFromUSec(totalValue, elapsedUSec); }

/* kind 0: latencyUsToHumanStr, 1: elapsedMSToHumanStr, 2: elapsedSecToHumanStr */
int64_t ref_unit_str(int kind, uint64_t value, char* outBuf, uint64_t outBufLen)
{
	switch(kind)
	{
		case 0: return copyOut(UnitTk::latencyUsToHumanStr(value), outBuf, outBufLen);
		case 1: return copyOut(UnitTk::elapsedMSToHumanStr(value), outBuf, outBufLen);
		case 2: return copyOut(UnitTk::elapsedSecToHumanStr(value), outBuf, outBufLen);
		default: return -1;
	}
}

/* returns 0 and *outBytes, or -1 and the exception text in outErr */
int ref_num_human_to_bytes(const char* numHuman, uint64_t* outBytes, char* outErr, uint64_t outErrLen)
{
	try
	{
		*outBytes = UnitTk::numHumanToBytesBinary(numHuman, true);
		return 0;
	}
	catch(ProgException& e)
	{
		copyOut(e.what(), outErr, outErrLen);
		return -1;
	}
}

/* HashTk::simple128 (service password hash, ProgArgs.cpp:2828) */
int64_t ref_simple128(const char* input, char* outBuf, uint64_t outBufLen)
	{ return copyOut(HashTk::simple128(input), outBuf, outBufLen); }

} // extern "C"
