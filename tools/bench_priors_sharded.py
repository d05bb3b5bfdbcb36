"""BASELINE configs[2]: per-character structure-prior generation only -- `--chars` random (label, w) pairs, characters sharded
over the ranks of one box, prior features all-gathered over NCCL for the SR decoder's concat (marconet_b200/parallel.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_priors_sharded.py --chars 1024 [--chunk 64] [--steps 3]

Prints one JSON line on rank 0: chars/s with and without the all-gather, max over ranks, CUDA events; and checks that
the gathered tensors equal an unsharded run on the first chunk.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chars", type=int, default=1024)
    ap.add_argument("--chunk", type=int, default=128, help="characters per generate_priors_sharded call (global)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--pipeline", type=int, default=1, help="sub-chunks per shard whose all-gather overlaps the next sub-chunk")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs left free for NCCL while the conv kernels run")
    args = ap.parse_args()
    from marconet_b200.models import networks
    from marconet_b200.parallel import generate_priors_sharded
    from marconet_b200.testing import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.reserve_sms > 0:
        from marconet_b200 import ops
        ops.set_max_ctas(torch.cuda.get_device_properties(local).multi_processor_count - args.reserve_sms)
    gen = networks.TSPGAN()
    gen.load_state_dict(synth.make_checkpoints(0)["tspgan"], strict=True)
    gen = gen.eval().to(dev)
    g = torch.Generator().manual_seed(11)
    labels = torch.randint(0, 6735, (args.chars, 1), generator=g).to(dev)
    styles = torch.randn(args.chars, 512, generator=g).to(dev)

    def run(gather):
        outs = []
        for c0 in range(0, args.chars, args.chunk):
            s, l = styles[c0:c0 + args.chunk], labels[c0:c0 + args.chunk]
            if gather:
                outs.append(generate_priors_sharded(gen, s, l, pipeline_chunks=args.pipeline))
            else:
                from marconet_b200.parallel import shard_range
                b, e = shard_range(s.shape[0], rank, world)
                outs.append(gen(s[b:e], l[b:e], None))
        return outs

    def timed(gather):
        for _ in range(2):
            run(gather)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            run(gather)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.no_grad():
        # correctness: gathered == unsharded on the first chunk
        full = gen(styles[:args.chunk], labels[:args.chunk], None)
        got = generate_priors_sharded(gen, styles[:args.chunk], labels[:args.chunk], pipeline_chunks=args.pipeline)
        err = max((a - b).abs().max().item() for a, b in zip(full, got))
        ms_nogather = timed(False)
        ms_gather = timed(True)
    if rank == 0:
        print(json.dumps({
            "metric": "prior_chars_per_sec", "n_gpus": world, "chars": args.chars, "chunk": args.chunk, "pipeline_chunks": args.pipeline, "reserve_sms": args.reserve_sms,
            "ms_no_gather": ms_nogather, "chars_per_s_no_gather": args.chars / ms_nogather * 1e3,
            "ms_with_allgather": ms_gather, "chars_per_s_with_allgather": args.chars / ms_gather * 1e3,
            "allgather_bytes_per_rank": int(args.chars * (256 * 64 * 64 + 512 * 32 * 32 + 3 * 128 * 128) * 4),
            "sharded_vs_unsharded_max_abs_diff": err}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
