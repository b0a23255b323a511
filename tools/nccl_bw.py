"""all_reduce bandwidth of the flat gradient size (80.9 MB fp32) between the visible GPUs: torchrun --nproc-per-node N tools/nccl_bw.py"""
import os, torch, torch.distributed as dist
local = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
for mb in (1, 18, 56, 81):
    x = torch.ones(mb * (1 << 20) // 4, device='cuda')
    for _ in range(5): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): dist.all_reduce(x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    if dist.get_rank() == 0:
        n = dist.get_world_size()
        print('all_reduce %3d MB over %d GPUs: %.3f ms, bus bw %.1f GB/s' % (mb, n, ms, mb / 1024 * 2 * (n - 1) / n / (ms * 1e-3)))
dist.destroy_process_group()
