import os, time, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
print("torch threads default", torch.get_num_threads())
a = torch.randn(8192, 256); b = torch.randn(256, 256)
for n in (4, 8, 16, 32, 64, 128):
    torch.set_num_threads(n)
    for _ in range(3): (a @ b)
    t = time.perf_counter()
    for _ in range(20): c = torch.nn.functional.softplus(a @ b, beta=100)
    print(n, "threads: %.2f ms per matmul+softplus" % ((time.perf_counter() - t) / 20 * 1e3))
