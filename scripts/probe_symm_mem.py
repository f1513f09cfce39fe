"""torchrun probe (2+ GPUs): what torch.distributed._symmetric_memory offers on this box - peer pointers, multicast
(NVLS) address, signal pads - the plumbing of the fused data-parallel optimizer step."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local)
    try:
        symm_mem.set_backend("CUDA")
    except Exception as e:
        print("set_backend:", repr(e))
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(rank + 1)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    info = dict(rank=rank, world=hdl.world_size, buffer_ptrs=[hex(p) for p in hdl.buffer_ptrs], multicast=hex(hdl.multicast_ptr or 0),
                has_multicast=getattr(hdl, "has_multicast_support", None), signal_pad_size=hdl.signal_pad_size,
                signal_pads=[hex(p) for p in hdl.signal_pad_ptrs], backend=symm_mem.get_backend(dev) if hasattr(symm_mem, "get_backend") else None)
    print(info, flush=True)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, "peer value", peer[:2].tolist(), flush=True)
    hdl.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
