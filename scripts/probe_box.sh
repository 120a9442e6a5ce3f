set -x
nproc; free -g | head -3; df -h / /tmp /dev/shm /root 2>&1; mount | grep -E ' / | /tmp | /dev/shm ' ; uname -r
nvidia-smi -L; nvidia-smi --query-gpu=name,memory.total,pcie.link.gen.current,pcie.link.width.current --format=csv
ls /usr/local/cuda/lib64 | grep -E 'cufile|nccl|curand' | head; ls /usr/local/cuda/gds 2>&1 | head; lsmod 2>/dev/null | grep -i nvidia | head; ls /etc/cufile.json 2>&1
python - <<'PY'
import os, ctypes, time, mmap
# O_DIRECT probe on /tmp, /dev/shm, repo
for d in ['/tmp','/dev/shm',os.getcwd()]:
    p=os.path.join(d,'odirect_probe.bin')
    try:
        fd=os.open(p,os.O_CREAT|os.O_RDWR|os.O_DIRECT,0o600)
        m=mmap.mmap(-1,1<<20)
        n=os.pwrite(fd,m,0)
        print(d,'O_DIRECT ok',n)
        os.close(fd)
    except Exception as e: print(d,'O_DIRECT fail',e)
    try: os.unlink(p)
    except: pass
# io_uring probe
libc=ctypes.CDLL(None,use_errno=True)
class P(ctypes.Structure): _fields_=[('x',ctypes.c_uint8*120)]
p=P(); r=libc.syscall(425,8,ctypes.byref(p)); print('io_uring_setup ->',r,ctypes.get_errno())
ctx=ctypes.c_ulong(0); r=libc.syscall(206,64,ctypes.byref(ctx)); print('io_setup ->',r,ctypes.get_errno())
# page-cache write throughput 1 thread
for d in ['/tmp','/dev/shm']:
    p=os.path.join(d,'tp.bin'); fd=os.open(p,os.O_CREAT|os.O_RDWR|os.O_TRUNC,0o600)
    b=bytes(1<<20); t=time.time()
    for i in range(2048): os.pwrite(fd,b,i<<20)
    dt=time.time()-t; print(d,'write GiB/s',2/dt)
    t=time.time()
    for i in range(2048): os.pread(fd,1<<20,i<<20)
    dt=time.time()-t; print(d,'read GiB/s',2/dt)
    os.close(fd); os.unlink(p)
PY
cat /proc/cpuinfo | grep 'model name' | sort | uniq -c
ulimit -l; cat /proc/meminfo | head -5
