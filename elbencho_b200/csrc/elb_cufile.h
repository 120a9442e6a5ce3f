/*
 * GPUDirect Storage access through libcufile, bound at run time (dlopen) so that the library has
 * no link-time dependency on it.
 *
 * Reference counterparts: CuFileHandleData (source/CuFileHandleData.h:20-77), cuFileDriverOpen in
 * ProgArgs (source/ProgArgs.cpp:2579), cuFileBufRegister in allocGPUIOBuffer
 * (source/workers/LocalWorker.cpp:1495-1509) and the cuFile wrappers (:2600-2713). The reference
 * only uses the synchronous calls with one block in flight and forbids --cufile with iodepth > 1
 * (ProgArgs.cpp:1312-1313); here the batch API (cuFileBatchIOSetUp/Submit/GetStatus) carries
 * iodepth > 1.
 *
 * The library searched is $ELB_CUFILE_LIB if set, else libcufile.so.0 / libcufile.so on the
 * loader path and /usr/local/cuda/lib64. If it cannot be loaded or a call fails, the worker fails
 * with the cuFile error text; there is no silent fallback to the staged path.
 */
#ifndef ELB_CUFILE_H_
#define ELB_CUFILE_H_

#include <cufile.h>

#include <mutex>
#include <string>

#include "elb_host.h"

namespace elb
{

class CuFileApi
{
	public:
		/* process-wide instance; loads the library on first use. @throw WorkerError */
		static CuFileApi& get();

		/* cuFileDriverOpen once per process (reference: ProgArgs.cpp:2579). @throw WorkerError */
		void driverOpenOnce();

		static std::string errorStr(CUfileError_t status);

		// function table
		CUfileError_t (*DriverOpen)(void);
		CUfileError_t (*DriverClose)(void);
		CUfileError_t (*HandleRegister)(CUfileHandle_t* fh, CUfileDescr_t* descr);
		void (*HandleDeregister)(CUfileHandle_t fh);
		CUfileError_t (*BufRegister)(const void* bufPtrBase, size_t length, int flags);
		CUfileError_t (*BufDeregister)(const void* bufPtrBase);
		ssize_t (*Read)(CUfileHandle_t fh, void* bufPtrBase, size_t size, off_t fileOffset,
			off_t bufPtrOffset);
		ssize_t (*Write)(CUfileHandle_t fh, const void* bufPtrBase, size_t size, off_t fileOffset,
			off_t bufPtrOffset);
		CUfileError_t (*BatchIOSetUp)(CUfileBatchHandle_t* batchIdp, unsigned nr);
		CUfileError_t (*BatchIOSubmit)(CUfileBatchHandle_t batchIdp, unsigned nr,
			CUfileIOParams_t* iocbp, unsigned int flags);
		CUfileError_t (*BatchIOGetStatus)(CUfileBatchHandle_t batchIdp, unsigned minNr,
			unsigned* nr, CUfileIOEvents_t* iocbp, struct timespec* timeout);
		void (*BatchIODestroy)(CUfileBatchHandle_t batchIdp);
		CUfileError_t (*BatchIOCancel)(CUfileBatchHandle_t batchIdp); // (may be NULL in old libs)

		const std::string& getLibPath() const { return libPath; }

	private:
		CuFileApi();

		void* libHandle{NULL};
		std::string libPath;
		std::mutex driverMutex;
		bool driverOpened{false};
};

/* registered cuFile handle of one file descriptor (reference: CuFileHandleData.h) */
class CuFileHandle
{
	public:
		CuFileHandle() {}
		~CuFileHandle() { deregister(); }

		CuFileHandle(const CuFileHandle&) = delete;
		CuFileHandle& operator=(const CuFileHandle&) = delete;

		/* @throw WorkerError with the cuFile error text */
		void registerFD(int fd, const std::string& pathForLog);
		void deregister();

		bool isRegistered() const { return registered; }
		CUfileHandle_t get() const { return handle; }

	private:
		CUfileHandle_t handle{NULL};
		bool registered{false};
};

} // namespace elb

#endif /* ELB_CUFILE_H_ */
