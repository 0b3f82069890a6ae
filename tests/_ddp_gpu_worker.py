"""Worker of tests/test_gpu_multi.py (one process per GPU under torch.distributed.run, NCCL).

Checks, on the real multi-GPU path of Trainer (flat arena, bucketed NCCL all-reduce overlapped with the backward pass,
1/world folded into the fused SGD kernel -- the replacement of DistributedDataParallel, trainer.py:79-82 of the reference):
  1. after construction every rank holds rank 0's parameters and BN buffers (the models are seeded per rank);
  2. one step: the reduced gradient arena equals the SUM of the per-rank local gradients (so SGD's 1/world gives the
     mean), bit for bit on every rank;
  3. after 3 more steps through Trainer.train (eager warm-up, CUDA-graph capture with the all-reduce inside, replay)
     the parameters are bit-identical on all ranks; BN running statistics are per rank by design;
  4. SyncBatchNorm (engine.enable_sync_batchnorm, --sync-bn): W ranks x B samples give the statistics, logits and
     (averaged) gradients of ONE process running the W*B batch.
Prints one line 'DDP_CHECK OK ...' on rank 0 or raises."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from convnet.pytorch_b200 import models
    from convnet.pytorch_b200.engine import convert_b200
    from convnet.pytorch_b200.trainer import Trainer
    from convnet.pytorch_b200.utils.optim import OptimRegime
    from convnet.pytorch_b200.utils.cross_entropy import CrossEntropyLoss

    torch.manual_seed(123 + rank)                      # different init per rank: the constructor must fix that
    model = models.resnet(dataset='imagenet', depth=18)
    convert_b200(model, dev)
    opt = OptimRegime(model, model.regime)
    tr = Trainer(model, CrossEntropyLoss().to(dev), opt, device_ids=[local], device=str(dev), distributed=True,
                 local_rank=local, print_freq=10 ** 9)
    arena = model._b200.arena

    def gather(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t.contiguous())
        return out

    # 1. broadcast at construction
    ps = gather(arena.p32)
    assert all(torch.equal(ps[0], p) for p in ps), 'parameters differ across ranks after Trainer construction'
    for name, buf in model.named_buffers():
        bs = gather(buf.float())
        assert all(torch.equal(bs[0], b) for b in bs), 'buffer %s differs after construction' % name

    g = torch.Generator().manual_seed(1000 + rank)     # per-rank data
    batches = [(torch.randn(32, 3, 64, 64, generator=g), torch.randint(0, 1000, (32,), generator=g)) for _ in range(4)]

    # 2. reduced gradients == sum of local gradients
    x, y = batches[0][0].to(dev), batches[0][1].to(dev)
    model.train()
    hook, model._b200.grad_bucket_hook = model._b200.grad_bucket_hook, None
    arena.zero_grad_force()
    model._b200.train_step(x, y, 0.0, None)            # local gradients only
    torch.cuda.synchronize()
    local_g = arena.g32.clone()
    want = torch.stack(gather(local_g)).sum(0)         # fixed summation order, same on every rank
    model._b200.grad_bucket_hook = hook
    # BN running statistics moved in the local pass; that is fine (per-rank buffers), parameters did not
    arena.zero_grad_force()
    model._b200.train_step(x, y, 0.0, None)            # same batch again, now with the bucketed all-reduce
    tr._allreduce_gradients()
    torch.cuda.synchronize()
    assert hook is None or hook.launched >= 2, 'expected several gradient buckets, saw %s' % getattr(hook, 'launched', 0)
    rel = float((arena.g32 - want).norm() / want.norm())
    if world == 2:
        assert torch.equal(arena.g32, want), 'reduced gradients != sum of local gradients (rel %.3e)' % rel
    else:
        assert rel < 1e-6, 'reduced gradients != sum of local gradients (rel %.3e)' % rel
    gs = gather(arena.g32)
    assert all(torch.equal(gs[0], t) for t in gs), 'reduced gradients differ across ranks'
    arena.zero_grad_force()

    # 3. training: parameters stay bit-identical
    res = tr.train(batches)
    torch.cuda.synchronize()
    ps = gather(arena.p32)
    assert all(torch.equal(ps[0], p) for p in ps), 'parameters diverged across ranks after %d steps' % tr.training_steps
    assert not torch.equal(ps[0], want.new_zeros(()).expand_as(ps[0])), 'parameters are zero'
    rm = gather(model.bn1.running_mean)
    per_rank_bn = not all(torch.equal(rm[0], t) for t in rm)
    # 4. SyncBatchNorm: this rank's 16 samples with synchronised statistics vs the concatenated batch in one process
    from convnet.pytorch_b200.engine import enable_sync_batchnorm
    torch.manual_seed(7)
    m_sync = convert_b200(models.resnet(dataset='imagenet', depth=18), dev)
    torch.manual_seed(7)
    m_full = convert_b200(models.resnet(dataset='imagenet', depth=18), dev)
    enable_sync_batchnorm(m_sync)
    gs = torch.Generator().manual_seed(555)
    xs = torch.randn(16 * world, 3, 64, 64, generator=gs).to(dev)
    ys = torch.randint(0, 1000, (16 * world,), generator=gs).to(dev)
    m_sync.train(); m_full.train()
    lo_s, _ = m_sync._b200.train_step(xs[16 * rank:16 * rank + 16], ys[16 * rank:16 * rank + 16], 0.0, None)
    g_sync = m_sync._b200.arena.g32.clone()
    dist.all_reduce(g_sync)
    g_sync /= world
    lo_f, _ = m_full._b200.train_step(xs, ys, 0.0, None)
    torch.cuda.synchronize()
    g_full = m_full._b200.arena.g32

    def rel(a, b):
        return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
    r_log = rel(lo_s, lo_f[16 * rank:16 * rank + 16])
    r_rm = max(rel(a, b) for (n, a), (_, b) in zip(m_sync.named_buffers(), m_full.named_buffers()) if 'running' in n)
    r_g = rel(g_sync, g_full)
    assert r_log < 2e-3 and r_rm < 1e-3 and r_g < 2e-2, 'SyncBatchNorm: logits %.3e running stats %.3e grads %.3e' % (
        r_log, r_rm, r_g)
    if rank == 0:
        print('DDP_CHECK OK world=%d steps=%d graph_replays=%d buckets_per_step=%s loss=%.4f per_rank_bn_stats=%s '
              'syncbn(logits %.2e stats %.2e grads %.2e)'
              % (world, tr.training_steps, tr.graph_replays, getattr(hook, 'launched', 0), res['loss'], per_rank_bn,
                 r_log, r_rm, r_g), flush=True)
    tr.release_graphs()
    dist.barrier()
    import threading
    t = threading.Timer(30.0, lambda: os._exit(0))     # a blocked teardown must not turn a passed check into a hang
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


if __name__ == '__main__':
    main()
