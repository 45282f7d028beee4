"""all-reduce time of the step's gradient buffers (49.2 MB fp32 table+MLP, 0.2 MB colour MLP) under NCCL. Measurement aid."""
import os, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
big = torch.zeros(12235152, device="cuda")
small = torch.zeros(55296, device="cuda")
def t(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
one = t(lambda: dist.all_reduce(big))
two = t(lambda: (dist.all_reduce(small), dist.all_reduce(big)))
q = big.numel() // 4
four = t(lambda: [dist.all_reduce(big[i * q:(i + 1) * q]) for i in range(4)])
half = t(lambda: dist.all_reduce(big[:big.numel() // 2]))
if dist.get_rank() == 0:
    print(f"[{os.environ.get('TAG','default')}] world {dist.get_world_size()}: 49MB {one:.1f} us ({2*big.numel()*4*(dist.get_world_size()-1)/dist.get_world_size()/one/1e3:.0f} GB/s bus), small+big {two:.1f} us, 4 quarters {four:.1f} us, half {half:.1f} us")
dist.destroy_process_group()
