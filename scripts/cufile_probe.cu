// Probe of libcufile on the GPU box: which API levels work (sync, batch, async), in which mode
// (GDS vs compat), on which filesystems. Exploration only; not part of the product.
#include <cuda_runtime.h>
#include <cufile.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }

int main(int argc, char** argv)
{
	const char* path = (argc > 1) ? argv[1] : "/dev/shm/cufile_probe.bin";
	const size_t block = 1 << 20;
	const size_t nblocks = 256;

	CUfileError_t st = cuFileDriverOpen();
	printf("cuFileDriverOpen: err=%d cu_err=%d\n", st.err, st.cu_err);

	CUfileDrvProps_t props;
	memset(&props, 0, sizeof(props));
	st = cuFileDriverGetProperties(&props);
	printf("props: err=%d nvfs major=%u minor=%u dstatusflags=0x%x dcontrolflags=0x%x max_direct_io=%zu KB\n",
		st.err, props.nvfs.major_version, props.nvfs.minor_version, props.nvfs.dstatusflags,
		props.nvfs.dcontrolflags, (size_t)props.nvfs.max_direct_io_size);

	int fd = open(path, O_CREAT | O_RDWR | O_DIRECT, 0644);
	printf("open(%s, O_DIRECT) fd=%d\n", path, fd);
	if(fd < 0) return 1;

	CUfileDescr_t descr;
	memset(&descr, 0, sizeof(descr));
	descr.handle.fd = fd;
	descr.type = CU_FILE_HANDLE_TYPE_OPAQUE_FD;
	CUfileHandle_t fh;
	st = cuFileHandleRegister(&fh, &descr);
	printf("cuFileHandleRegister: err=%d (%s)\n", st.err, CUFILE_ERRSTR(st.err));
	if(st.err != CU_FILE_SUCCESS) return 2;

	char* dev;
	cudaMalloc(&dev, block * nblocks);
	cudaMemset(dev, 0x5a, block * nblocks);
	st = cuFileBufRegister(dev, block * nblocks, 0);
	printf("cuFileBufRegister: err=%d (%s)\n", st.err, CUFILE_ERRSTR(st.err));

	double t0 = now();
	for(size_t i = 0; i < nblocks; i++)
	{
		ssize_t res = cuFileWrite(fh, dev, block, i * block, i * block);
		if(res != (ssize_t)block) { printf("cuFileWrite res=%zd\n", res); break; }
	}
	double t1 = now();
	printf("sync cuFileWrite 1MiB x %zu: %.2f GiB/s\n", nblocks, nblocks * block / (t1 - t0) / (1 << 30));

	t0 = now();
	for(size_t i = 0; i < nblocks; i++)
	{
		ssize_t res = cuFileRead(fh, dev, block, i * block, i * block);
		if(res != (ssize_t)block) { printf("cuFileRead res=%zd\n", res); break; }
	}
	t1 = now();
	printf("sync cuFileRead 1MiB x %zu: %.2f GiB/s\n", nblocks, nblocks * block / (t1 - t0) / (1 << 30));

	// batch: 64 x 4 KiB random-ish reads
	const unsigned nr = 64;
	CUfileBatchHandle_t batch;
	st = cuFileBatchIOSetUp(&batch, nr);
	printf("cuFileBatchIOSetUp: err=%d (%s)\n", st.err, CUFILE_ERRSTR(st.err));
	if(st.err == CU_FILE_SUCCESS)
	{
		CUfileIOParams_t params[nr];
		CUfileIOEvents_t events[nr];
		int rounds = 200;
		t0 = now();
		size_t done = 0;
		for(int r = 0; r < rounds; r++)
		{
			for(unsigned i = 0; i < nr; i++)
			{
				memset(&params[i], 0, sizeof(params[i]) );
				params[i].mode = CUFILE_BATCH;
				params[i].fh = fh;
				params[i].opcode = CUFILE_READ;
				params[i].u.batch.devPtr_base = dev;
				params[i].u.batch.devPtr_offset = i * 4096;
				params[i].u.batch.file_offset = ( (size_t)(i * 7919 + r * 13) % (nblocks * 256) ) * 4096;
				params[i].u.batch.size = 4096;
				params[i].cookie = (void*)(uintptr_t)i;
			}
			st = cuFileBatchIOSubmit(batch, nr, params, 0);
			if(st.err != CU_FILE_SUCCESS) { printf("cuFileBatchIOSubmit: err=%d (%s)\n", st.err, CUFILE_ERRSTR(st.err)); break; }
			unsigned got = 0;
			while(got < nr)
			{
				unsigned n = nr;
				struct timespec to = {1, 0};
				st = cuFileBatchIOGetStatus(batch, 1, &n, events, &to);
				if(st.err != CU_FILE_SUCCESS) { printf("cuFileBatchIOGetStatus: err=%d\n", st.err); got = nr; break; }
				for(unsigned k = 0; k < n; k++)
					if(events[k].status != CUFILE_COMPLETE || events[k].ret != 4096)
						{ printf("event status=%d ret=%zu\n", events[k].status, events[k].ret); }
				got += n;
			}
			done += nr;
		}
		t1 = now();
		printf("batch 64 x 4KiB reads: %.0f IOPS\n", done / (t1 - t0));
		cuFileBatchIODestroy(batch);
	}

	// async (stream ordered)
	cudaStream_t stream;
	cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
	st = cuFileStreamRegister(stream, 0xf);
	printf("cuFileStreamRegister: err=%d (%s)\n", st.err, CUFILE_ERRSTR(st.err));
	{
		size_t size = block; off_t foff = 0, boff = 0; ssize_t bytesRead = 0;
		st = cuFileReadAsync(fh, dev, &size, &foff, &boff, &bytesRead, stream);
		printf("cuFileReadAsync: err=%d (%s)\n", st.err, CUFILE_ERRSTR(st.err));
		cudaError_t ce = cudaStreamSynchronize(stream);
		printf("async read result: cuda=%d bytes=%zd\n", (int)ce, bytesRead);
	}
	cuFileStreamDeregister(stream);

	cuFileBufDeregister(dev);
	cuFileHandleDeregister(fh);
	close(fd);
	unlink(path);
	cuFileDriverClose();
	return 0;
}
