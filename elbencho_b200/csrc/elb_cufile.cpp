/*
 * Run-time binding of libcufile (see elb_cufile.h).
 */
#include <dlfcn.h>
#include <stdlib.h>

#include "elb_cufile.h"

namespace elb
{

template<typename FuncPtr>
static void loadSymbol(void* libHandle, const char* name, FuncPtr& outFunc,
	const std::string& libPath)
{
	void* symbol = dlsym(libHandle, name);

	if(!symbol)
		throw WorkerError(std::string("cuFile API requested, but symbol is missing in library. ") +
			"Library: " + libPath + "; Symbol: " + name);

	outFunc = reinterpret_cast<FuncPtr>(symbol);
}

CuFileApi::CuFileApi()
{
	std::vector<std::string> candidates;

	const char* envLib = getenv("ELB_CUFILE_LIB");
	if(envLib && envLib[0] )
		candidates.push_back(envLib);
	else
	{
		candidates.push_back("libcufile.so.0");
		candidates.push_back("libcufile.so");
		candidates.push_back("/usr/local/cuda/lib64/libcufile.so.0");
		candidates.push_back("/usr/local/cuda/targets/x86_64-linux/lib/libcufile.so.0");
	}

	std::string dlErrors;

	for(const std::string& candidate : candidates)
	{
		libHandle = dlopen(candidate.c_str(), RTLD_NOW | RTLD_LOCAL);

		if(libHandle)
		{
			libPath = candidate;
			break;
		}

		dlErrors += std::string(dlErrors.empty() ? "" : " | ") + dlerror();
	}

	if(!libHandle)
		throw WorkerError("cuFile API requested, but libcufile could not be loaded "
			"(set ELB_CUFILE_LIB to its path). Loader errors: " + dlErrors);

	loadSymbol(libHandle, "cuFileDriverOpen", DriverOpen, libPath);
	// (cufile.h maps cuFileDriverClose to cuFileDriverClose_v2 by macro)
	loadSymbol(libHandle, dlsym(libHandle, "cuFileDriverClose_v2") ?
		"cuFileDriverClose_v2" : "cuFileDriverClose", DriverClose, libPath);
	loadSymbol(libHandle, "cuFileHandleRegister", HandleRegister, libPath);
	loadSymbol(libHandle, "cuFileHandleDeregister", HandleDeregister, libPath);
	loadSymbol(libHandle, "cuFileBufRegister", BufRegister, libPath);
	loadSymbol(libHandle, "cuFileBufDeregister", BufDeregister, libPath);
	loadSymbol(libHandle, "cuFileRead", Read, libPath);
	loadSymbol(libHandle, "cuFileWrite", Write, libPath);
	loadSymbol(libHandle, "cuFileBatchIOSetUp", BatchIOSetUp, libPath);
	loadSymbol(libHandle, "cuFileBatchIOSubmit", BatchIOSubmit, libPath);
	loadSymbol(libHandle, "cuFileBatchIOGetStatus", BatchIOGetStatus, libPath);
	loadSymbol(libHandle, "cuFileBatchIODestroy", BatchIODestroy, libPath);
	BatchIOCancel = (CUfileError_t (*)(CUfileBatchHandle_t) )dlsym(libHandle, "cuFileBatchIOCancel");
}

CuFileApi& CuFileApi::get()
{
	static std::mutex instanceMutex;
	static CuFileApi* instance = NULL;

	std::unique_lock<std::mutex> lock(instanceMutex);

	if(!instance)
		instance = new CuFileApi(); // (throws on failure, so a later call retries)

	return *instance;
}

void CuFileApi::driverOpenOnce()
{
	std::unique_lock<std::mutex> lock(driverMutex);

	if(driverOpened)
		return;

	CUfileError_t openRes = DriverOpen();

	if(openRes.err != CU_FILE_SUCCESS)
		throw WorkerError("cuFile driver init failed (cuFileDriverOpen). "
			"cuFile Error: " + errorStr(openRes) );

	driverOpened = true;
}

std::string CuFileApi::errorStr(CUfileError_t status)
{
	std::string text = CUFILE_ERRSTR(status.err);

	if(status.err == CU_FILE_CUDA_DRIVER_ERROR)
		text += " (CUDA driver error " + std::to_string( (int)status.cu_err) + ")";

	return text + " [" + std::to_string( (int)status.err) + "]";
}

/* reference: CuFileHandleData::registerHandle (CuFileHandleData.h:33-55) */
void CuFileHandle::registerFD(int fd, const std::string& pathForLog)
{
	deregister();

	CuFileApi& api = CuFileApi::get();

	CUfileDescr_t descr;
	memset(&descr, 0, sizeof(descr) );
	descr.handle.fd = fd;
	descr.type = CU_FILE_HANDLE_TYPE_OPAQUE_FD;

	CUfileError_t registerRes = api.HandleRegister(&handle, &descr);

	if(registerRes.err != CU_FILE_SUCCESS)
		throw WorkerError("cuFile file handle registration failed (cuFileHandleRegister). "
			"Path: " + pathForLog + "; "
			"FD: " + std::to_string(fd) + "; "
			"cuFile Error: " + CuFileApi::errorStr(registerRes) );

	registered = true;
}

/* reference: CuFileHandleData::deregisterHandle (CuFileHandleData.h:62-72) */
void CuFileHandle::deregister()
{
	if(!registered)
		return;

	CuFileApi::get().HandleDeregister(handle);
	registered = false;
}

} // namespace elb
