"""Training / evaluation loop with the reference's ``Trainer`` interface (trainer.py:54-285).

    Trainer(model, criterion, optimizer=None, device_ids=[0], device='cuda', dtype=torch.float,
            distributed=False, local_rank=-1, adapt_grad_norm=None, mixup=None, cutmix=None,
            loss_scale=1., grad_clip=-1, print_freq=100)
    .train(loader, average_output=False, chunk_batch=1) / .validate(loader) / .calibrate_bn(loader, num_steps)
        -> dict of meter averages (step, data, loss, prec1, prec5, error1, error5[, grad])

One iteration = zero_grad -> regime update -> H2D -> forward -> loss -> backward -> unscale -> clip ->
optimizer step, as in ``Trainer._step`` (trainer.py:106-177).

B200 path (model converted by engine.convert_b200): inputs stay fp32 NCHW on the device (the stem kernel
does the bf16/NHWC conversion), gradients land in the flat arena, the data-parallel reduction is ONE
NCCL all-reduce of that arena (no DistributedDataParallel wrapper, no per-forward buffer broadcast --
SURVEY.md section 2.3 C1/C2), and the per-tensor unscale / clip loops of the reference
(trainer.py:165-172) are folded into the fused optimizer kernel.
"""
import logging
import os
import time

import torch
import torch.nn as nn
import torch.distributed as dist
from torch.nn.utils import clip_grad_norm_

from .utils import regularization
from .utils.meters import AverageMeter, accuracy

_METERS = ('step', 'data', 'loss', 'prec1', 'prec5')


def _flatten_duplicates(inputs, target, batch_first=True, expand_target=True):
    """[B, D, C, H, W] batch-augmentation input -> [(B*D), C, H, W]; targets repeated to match."""
    copies = inputs.size(1)
    if not batch_first:
        inputs = inputs.transpose(0, 1)
    inputs = inputs.flatten(0, 1)
    if expand_target:
        if batch_first:
            target = target.view(-1, 1).expand(-1, copies)
        else:
            target = target.view(1, -1).expand(copies, -1)
        target = target.flatten(0, 1)
    return inputs, target


def _average_duplicates(outputs, target, batch_first=True):
    """Mean of the network outputs over the duplicates of each sample (target is NOT expanded)."""
    bsz = target.size(0)
    if batch_first:
        return outputs.view(bsz, -1, *outputs.shape[1:]).mean(dim=1)
    return outputs.view(-1, bsz, *outputs.shape[1:]).mean(dim=0)


def _cuda_prefetch(loader, device, dtype):
    """Yield device-resident batches one step ahead: batch i+1 is copied host->device on a side stream while
    step i computes (the reference issues a blocking copy at the top of every step, trainer.py:116-117).
    The copies land in a ring of three persistent device buffers per (shape, dtype): no caching-allocator traffic per
    step (a fresh 154 MB tensor per batch that is handed across streams made the allocator stall now and then)."""
    copy_stream = torch.cuda.Stream(device=device)
    ring = {}          # (shape, dtype) of x and y -> [[x_buf, y_buf, consumed_event or None], ...]
    turn = {}

    def stage(inputs, target):
        x_dtype = inputs.dtype if inputs.dtype == torch.uint8 else dtype     # uint8 image batches stay uint8
        key = (tuple(inputs.shape), x_dtype, tuple(target.shape), target.dtype)
        slots = ring.setdefault(key, [])
        k = turn.get(key, 0)
        turn[key] = (k + 1) % 3
        if len(slots) <= k:
            slots.append([torch.empty(inputs.shape, device=device, dtype=x_dtype),
                          torch.empty(target.shape, device=device, dtype=target.dtype), None])
        slot = slots[k]
        with torch.cuda.stream(copy_stream):
            if slot[2] is not None:
                copy_stream.wait_event(slot[2])            # the step that read this slot has been enqueued AND has run
            slot[0].copy_(inputs, non_blocking=True)
            slot[1].copy_(target, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        return slot, ready

    def hand_over(slot, ready):
        torch.cuda.current_stream(device).wait_event(ready)
        return slot[0], slot[1]

    def release(slot):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))       # everything the consumer enqueued on this batch
        slot[2] = ev

    pending = None
    for inputs, target in loader:
        nxt = stage(inputs, target)
        if pending is not None:
            yield hand_over(*pending)
            release(pending[0])
        pending = nxt
    if pending is not None:
        yield hand_over(*pending)
        release(pending[0])


class Trainer(object):
    def __init__(self, model, criterion, optimizer=None, device_ids=[0], device='cuda', dtype=torch.float,
                 distributed=False, local_rank=-1, adapt_grad_norm=None, mixup=None, cutmix=None,
                 loss_scale=1., grad_clip=-1, print_freq=100):
        if mixup is not None or cutmix is not None:
            raise NotImplementedError('mixup / cutmix are outside the B200 hot path (SURVEY.md section 2, #18)')
        self._model = model
        self.criterion = criterion
        self.epoch = 0
        self.training_steps = 0
        self.optimizer = optimizer
        self.device = device
        self.dtype = dtype
        self.distributed = distributed
        self.local_rank = local_rank
        self.print_freq = print_freq
        self.grad_clip = grad_clip
        self.grad_scale = None
        self.loss_scale = loss_scale
        self.adapt_grad_norm = adapt_grad_norm
        self.b200 = getattr(model, '_b200', None)
        self._graphs, self._graph_pool, self._graph_broken, self._graph_static_ok = {}, None, False, None
        self.use_graphs = os.environ.get('B200_CUDA_GRAPH', '1') != '0'
        self.graph_replays = 0                 # bench.py: launches replayed from graphs are not counted by the library
        self.graph_replayed_launches = 0
        self.world_size = dist.get_world_size() if (distributed and dist.is_initialized()) else 1
        self._upstream_t, self._upstream_v = None, None

        if self.b200 is not None:
            self.model = model
            if optimizer is not None and hasattr(optimizer, 'fold_zero_grad'):
                optimizer.fold_zero_grad(True)     # the fused SGD pass also clears the gradient arena
            if distributed and self.world_size > 1:
                self._broadcast_initial_state()
                if os.environ.get('B200_AR_OVERLAP', '1') != '0':
                    self.b200.grad_bucket_hook = Trainer._GradBuckets(self.b200.arena, self.b200.device)
        elif distributed:
            if device_ids and 'cuda' in str(device):
                self.model = nn.parallel.DistributedDataParallel(model, device_ids=device_ids,
                                                                 output_device=device_ids[0])
            else:  # CPU / gloo processes (the reference crashes here: trainer.py:82 with device_ids=None)
                self.model = nn.parallel.DistributedDataParallel(model)
        elif device_ids and len(device_ids) > 1:
            self.model = nn.DataParallel(model, device_ids)
        else:
            self.model = model

    # ------------------------------------------------------------------ data-parallel helpers (B200)
    def _broadcast_initial_state(self):
        """rank 0 -> all: parameters (one flat broadcast) and BN buffers, once (what DDP's ctor does)."""
        arena = self.b200.arena
        dist.broadcast(arena.p32, src=0)
        for buf in self._model.buffers():
            dist.broadcast(buf, src=0)
        arena.sync_shadow()

    def _allreduce_gradients(self):
        """Sum of the gradient arena over the ranks (the 1/world factor is folded into the SGD kernel).  With the
        bucketed path (default) the reduction already ran inside the backward pass -- NCCL all-reduces of arena ranges
        on a communication stream, launched as soon as a range is final and overlapped with the remaining backward
        kernels (inside the captured graph they are graph nodes) -- and nothing is left to do here."""
        if self.b200 is None or self.world_size <= 1 or self.b200.grad_bucket_hook is not None:
            return
        from . import ops
        g32 = self.b200.arena.g32
        with ops._T('allreduce_nccl', 0, 4 * g32.numel()):
            dist.all_reduce(g32)

    class _GradBuckets(object):
        """engine.Runtime.grad_bucket_hook for data-parallel runs (replaces DistributedDataParallel's bucketed reducer,
        trainer.py:79-82 of the reference)."""

        def __init__(self, arena, device):
            self.g32 = arena.g32
            # host tensors (gloo, CPU tests of the N > 1 logic) have no streams: the reduction is then synchronous
            self.comm = torch.cuda.Stream(device=device) if torch.device(device).type == 'cuda' else None
            self.launched = 0
            self.bytes = 0

        def bucket(self, lo, hi, wg_stream):
            from . import ops
            seg = self.g32[lo:hi]
            self.launched += 1
            self.bytes += 4 * (hi - lo)
            if self.comm is None:
                dist.all_reduce(seg)
                return
            main = torch.cuda.current_stream()
            if ops._TIMING is not None:            # bench.py's per-class timing step: serialised on the main stream
                if wg_stream is not None:
                    main.wait_stream(wg_stream)
                with ops._T('allreduce_nccl', 0, 4 * seg.numel()):
                    dist.all_reduce(seg)
                return
            ev = torch.cuda.Event()
            ev.record(main)
            self.comm.wait_event(ev)
            if wg_stream is not None:              # the weight gradients of this range were produced on the side stream
                ev2 = torch.cuda.Event()
                ev2.record(wg_stream)
                self.comm.wait_event(ev2)
            with torch.cuda.stream(self.comm):
                dist.all_reduce(seg)

        def finish(self):
            if self.comm is not None:
                torch.cuda.current_stream().wait_stream(self.comm)

    # ------------------------------------------------------------------ step capture (B200)
    # The forward + loss + backward of one batch is ~550 kernel launches; issued eagerly from Python they cost
    # about as much host time as the GPU needs to run them.  After two eager steps per (shape, scale) key the
    # sequence is captured once into a CUDA graph and replayed.  The optimiser
    # update and the gradient all-reduce stay outside the graph, so learning-rate schedules keep working.
    def _graph_eligible(self):
        if self.b200 is None or not self.use_graphs or self._graph_broken:
            return False
        return self._hooks_static()

    def _hooks_static(self):
        """True when no regularizer needs to run between forward and backward (such hooks can neither be replayed from
        a graph nor wrapped around the fused forward+loss+backward call)."""
        if self._graph_static_ok is None:
            ok = True    # dropout is fine: the mask comes from torch's graph-safe CUDA generator (philox offsets advance per replay)
            opt = self.optimizer
            for o in getattr(opt, 'optim_regime_list', [opt]):
                reg = getattr(o, 'regularizer', None)
                for r in getattr(reg, 'regularization_list', []):
                    if type(r).pre_forward is not regularization.Regularizer.pre_forward or \
                            type(r).pre_backward is not regularization.Regularizer.pre_backward:
                        ok = False
            self._graph_static_ok = ok
        return self._graph_static_ok

    def graphed_forward_backward(self, inputs, target):
        """Forward + criterion + backward of one device-resident batch through a captured CUDA graph.
        Returns (logits, loss, stats) -- detached device tensors; stats = fp32[3] {loss, top-1 %, top-5 %} when the fused
        loss kernel computed them, else None -- or None when this call has to run eagerly (warm-up steps of
        a new shape, unsupported configuration).  Gradients land in the arena exactly as in the eager path."""
        if not self._graph_eligible() or not inputs.is_cuda:
            return None
        # loss / gradient scales are NOT part of the key: they reach the kernels through a device scalar
        key = (tuple(inputs.shape), inputs.dtype, tuple(target.shape), target.dtype, self._model.training)
        st = self._graphs.get(key)
        if st is None:
            st = self._graphs[key] = {'seen': 0, 'graph': None}
        st['seen'] += 1
        if st['graph'] is None:
            if st['seen'] <= 2:
                return None                       # eager warm-up (library handles, allocator, autotuned state)
            try:
                self._capture(st, inputs, target)
            except Exception as e:  # noqa: BLE001  -- keep training eagerly if capture is impossible here
                if self.b200.grad_bucket_hook is not None:
                    # NCCL inside the capture is the likely culprit: fall back to ONE flat all-reduce after the graph
                    logging.warning('B200: capture with in-graph all-reduce failed (%s); retrying without overlap', e)
                    self.b200.grad_bucket_hook = None
                    self.b200.arena.zero_grad_force()
                    st['seen'] = 2
                    return None
                logging.warning('B200: CUDA-graph capture failed (%s); continuing with eager launches', e)
                self._graph_broken = True
                self._graphs.clear()
                return None
        st['x'].copy_(inputs, non_blocking=True)
        st['y'].copy_(target, non_blocking=True)
        self._upstream()                          # refresh the device scalar if a scale changed
        st['graph'].replay()
        self.graph_replays += 1
        self.graph_replayed_launches += st['launches']
        return st['out'].detach(), st['loss'].detach(), st['stats']

    def _capture(self, st, inputs, target):
        from . import lib
        x_s, y_s = torch.empty_like(inputs), torch.empty_like(target)
        x_s.copy_(inputs)
        y_s.copy_(target)
        if self._graph_pool is None:
            self._graph_pool = torch.cuda.graph_pool_handle()
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        n0 = lib.launch_count()
        up = self._upstream()
        eps = self._plain_ce_eps()
        # with NCCL all-reduces inside the capture, ProcessGroupNCCL's watchdog thread polls CUDA events concurrently: the
        # default "global" capture mode would treat that as a capture violation
        mode = 'thread_local' if self.b200.grad_bucket_hook is not None else 'global'
        # capture on a HIGH-priority stream: the kernel nodes inherit it, so when a main-chain kernel (BN backward ->
        # dgrad of the next unit) and a side-stream weight gradient become ready together the block scheduler starts the
        # critical one first and the wgrad CTAs fill in beside the HBM-bound BN kernels that follow
        if getattr(self, '_capture_stream', None) is None:
            prio = -1 if os.environ.get('B200_MAIN_PRIORITY', '1') != '0' else 0
            self._capture_stream = torch.cuda.Stream(device=x_s.device, priority=prio)
        with torch.cuda.graph(graph, pool=self._graph_pool, stream=self._capture_stream, capture_error_mode=mode):
            stats = None
            if eps is not None:            # the whole step is library calls: nothing of autograd inside the graph
                out, stats = self.b200.train_step(x_s, y_s, eps, up)
                loss = stats[0]
            else:
                out = self.model(x_s)
                loss = self.criterion(out, y_s)
                torch.autograd.backward(loss, grad_tensors=[up])
        st.update(graph=graph, x=x_s, y=y_s, out=out, loss=loss, stats=stats, launches=lib.launch_count() - n0)

    def release_graphs(self):
        """Drop every captured step (call before tearing the process group down: graphs that captured NCCL all-reduces
        keep the communicator busy and destroy_process_group() can block on them)."""
        self._graphs.clear()
        self._graph_pool = None
        import gc
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def _plain_ce_eps(self):
        """label-smoothing coefficient when the criterion is the reference's plain CrossEntropyLoss (class indices,
        mean reduction, no weights / soft targets / ignore index) -- the case engine.Runtime.train_step fuses;
        None otherwise (generic autograd path)."""
        from .utils.cross_entropy import CrossEntropyLoss
        c = self.criterion
        if type(c) is CrossEntropyLoss and c.weight is None and c.smooth_dist is None and c.reduction == 'mean' \
                and c.ignore_index < 0 and c.from_logits:
            return float(c.smooth_eps or 0.0)
        return None

    def _upstream(self):
        """d(scaled loss)/d(loss) = grad_scale * loss_scale (trainer.py:158-161 of the reference multiplies the loss)
        as a persistent 0-dim device tensor handed to autograd.backward: no per-step scalar kernels, and a captured
        graph reads the current value instead of a baked-in constant."""
        v = 1.0
        if self.grad_scale is not None:
            v *= float(self.grad_scale)
        if self.loss_scale is not None:
            v *= float(self.loss_scale)
        if self._upstream_t is None:
            self._upstream_t = torch.empty((), device=self.b200.device, dtype=torch.float32)
        if v != self._upstream_v:
            self._upstream_t.fill_(v)
            self._upstream_v = v
        return self._upstream_t

    # ------------------------------------------------------------------ one optimisation step
    def _input_dtype(self):
        return torch.float if self.b200 is not None else self.dtype

    def _grad_norm(self, inputs_batch, target_batch, chunk_batch=1):
        if self.b200 is not None:
            # nn.Module.zero_grad() only drops the views (p.grad = None): the kernels accumulate into the arena itself
            self.b200.arena.zero_grad()
            self.b200.arena.rebind_grads()
        else:
            self.model.zero_grad()
        for inputs, target in zip(inputs_batch.chunk(chunk_batch, dim=0), target_batch.chunk(chunk_batch, dim=0)):
            target = target.to(self.device)
            inputs = inputs.to(self.device, dtype=self._input_dtype())
            loss = self.criterion(self.model(inputs), target)
            if chunk_batch > 1:
                loss = loss / chunk_batch
            loss.backward()
        return clip_grad_norm_(self.model.parameters(), float('inf'))

    def _step(self, inputs_batch, target_batch, training=False, average_output=False, chunk_batch=1):
        """-> (outputs, loss, grad norm or None).  ``loss`` is a float, or -- on the fused B200 path -- the fp32[3] device
        tensor {loss, top-1 %, top-5 %} of the loss kernel, which Trainer.forward reads back asynchronously
        (the reference synchronises three times per step here: trainer.py:153,226-227)."""
        outputs, total_loss, grad, stats = [], 0, None, None
        if training:
            self.optimizer.zero_grad()
            self.optimizer.update(self.epoch, self.training_steps)

        chunks = zip(inputs_batch.chunk(chunk_batch, dim=0), target_batch.chunk(chunk_batch, dim=0))
        for i, (inputs, target) in enumerate(chunks):
            target = target.to(self.device, non_blocking=True)
            if self.b200 is not None and inputs.dtype == torch.uint8:
                inputs = inputs.to(self.device, non_blocking=True)
            else:
                inputs = inputs.to(self.device, dtype=self._input_dtype(), non_blocking=True)
            if training and chunk_batch == 1 and not average_output:
                replayed = self.graphed_forward_backward(inputs, target)
                if replayed is not None:
                    outputs.append(replayed[0])
                    if replayed[2] is not None:
                        stats = replayed[2]
                    else:
                        total_loss += float(replayed[1])
                    continue
            if training and self.b200 is not None and chunk_batch == 1 and not average_output and inputs.is_cuda \
                    and self._hooks_static() and self._plain_ce_eps() is not None \
                    and target.dtype == torch.long and target.dim() == 1:
                # eager form of the captured step (warm-up iterations of a new shape, B200_CUDA_GRAPH=0)
                self.optimizer.pre_forward()
                output, stats = self.b200.train_step(inputs, target, self._plain_ce_eps(), self._upstream())
                outputs.append(output)
                self.optimizer.pre_backward()
                continue
            if training:
                self.optimizer.pre_forward()
            output = self.model(inputs)
            if average_output:
                if isinstance(output, (list, tuple)):
                    output = [_average_duplicates(o, target) if o is not None else None for o in output]
                else:
                    output = _average_duplicates(output, target)
            loss = self.criterion(output, target)
            if chunk_batch > 1:
                loss = loss / chunk_batch
            if isinstance(output, (list, tuple)):
                output = output[0]
            outputs.append(output.detach())
            total_loss += float(loss.detach())

            if training:
                if i == 0:
                    self.optimizer.pre_backward()
                if self.b200 is not None and loss.dim() == 0 and loss.dtype == torch.float32:
                    torch.autograd.backward(loss, grad_tensors=[self._upstream()])
                else:
                    if self.grad_scale is not None:
                        loss = loss * self.grad_scale
                    if self.loss_scale is not None:
                        loss = loss * self.loss_scale
                    loss.backward()

        if training:
            if self.b200 is not None:
                self._allreduce_gradients()
                self.optimizer.set_grad_unscale(self.loss_scale if self.loss_scale is not None else 1.0,
                                                self.world_size)
                if self.grad_clip > 0:
                    self.optimizer.request_clip(self.grad_clip)
                self.optimizer.step()
                if self.grad_clip > 0:
                    grad = self.optimizer.last_grad_norm
            else:
                if self.loss_scale is not None:
                    for p in self.model.parameters():
                        if p.grad is not None:
                            p.grad.data.div_(self.loss_scale)
                if self.grad_clip > 0:
                    grad = clip_grad_norm_(self.model.parameters(), self.grad_clip)
                self.optimizer.step()
            self.training_steps += 1

        return (outputs[0] if len(outputs) == 1 else torch.cat(outputs, dim=0)), (stats if stats is not None else total_loss), grad

    # ------------------------------------------------------------------ epoch loop
    def forward(self, data_loader, num_steps=None, training=False, average_output=False, chunk_batch=1):
        meters = {name: AverageMeter() for name in _METERS}
        if training and self.grad_clip > 0:
            meters['grad'] = AverageMeter()
        batch_first = not ((training and isinstance(self.model, nn.DataParallel)) or chunk_batch > 1)

        def summary():
            res = {name: m.avg for name, m in meters.items()}
            res['error1'] = 100. - res['prec1']
            res['error5'] = 100. - res['prec5']
            return res

        try:
            n_batches = len(data_loader)
        except TypeError:
            n_batches = -1
        tick = time.time()
        batches = _cuda_prefetch(data_loader, self.device, self._input_dtype()) \
            if (self.b200 is not None and chunk_batch == 1) else data_loader
        # lazy meters (B200 fused path): the step's {loss, prec1, prec5} come from the loss kernel as one device
        # tensor; it is copied to pinned host memory asynchronously every step and only awaited when a log line is
        # due, the ring is full or the loop ends -- no host synchronisation inside a step
        pending, ring = [], 8
        pinned = None

        def drain(keep=0):
            while len(pending) > keep:
                slot, ev, n = pending.pop(0)
                ev.synchronize()
                v = pinned[slot]
                meters['loss'].update(float(v[0]), n)
                meters['prec1'].update(float(v[1]), n)
                meters['prec5'].update(float(v[2]), n)

        for i, (inputs, target) in enumerate(batches):
            duplicates = inputs.dim() > 4  # B x D x C x H x W
            if training and duplicates and self.adapt_grad_norm is not None and i % self.adapt_grad_norm == 0:
                per_copy = sum(float(self._grad_norm(inputs.select(1, j), target)) for j in range(inputs.size(1)))
                per_copy /= inputs.size(1)
                joint = float(self._grad_norm(*_flatten_duplicates(inputs, target, batch_first)))
                self.grad_scale = per_copy / joint
                logging.info('New loss scale: %s', self.grad_scale)

            meters['data'].update(time.time() - tick)
            if duplicates:
                inputs, target = _flatten_duplicates(inputs, target, batch_first,
                                                     expand_target=not average_output)
            output, loss, grad = self._step(inputs, target, training=training, average_output=average_output,
                                            chunk_batch=chunk_batch)
            n = inputs.size(0)
            if torch.is_tensor(loss):              # fused statistics: asynchronous read-back
                if pinned is None:
                    pinned = torch.empty((ring, 3), dtype=torch.float32).pin_memory()
                drain(keep=ring - 1)
                slot = i % ring
                pinned[slot].copy_(loss, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pending.append((slot, ev, n))
                if i % self.print_freq == 0 or i == n_batches - 1:
                    drain()
            else:
                drain()
                target_dev = target.to(output.device)
                prec1, prec5 = accuracy(output, target_dev, topk=(1, 5))
                meters['loss'].update(float(loss), n)
                meters['prec1'].update(float(prec1), n)
                meters['prec5'].update(float(prec5), n)
            if grad is not None:
                meters['grad'].update(float(grad), n)
            meters['step'].update(time.time() - tick)
            tick = time.time()

            if i % self.print_freq == 0 or i == n_batches - 1:
                msg = ('{phase} - Epoch: [{0}][{1}/{2}]\t'
                       'Time {m[step].val:.3f} ({m[step].avg:.3f})\t'
                       'Data {m[data].val:.3f} ({m[data].avg:.3f})\t'
                       'Loss {m[loss].val:.4f} ({m[loss].avg:.4f})\t'
                       'Prec@1 {m[prec1].val:.3f} ({m[prec1].avg:.3f})\t'
                       'Prec@5 {m[prec5].val:.3f} ({m[prec5].avg:.3f})\t').format(
                    self.epoch, i, n_batches, phase='TRAINING' if training else 'EVALUATING', m=meters)
                if 'grad' in meters:
                    msg += 'Grad {m[grad].val:.3f} ({m[grad].avg:.3f})'.format(m=meters)
                logging.info(msg)
            if num_steps is not None and i >= num_steps:  # (sic) the reference runs num_steps+1 iterations
                break
        drain()
        return summary()

    def train(self, data_loader, average_output=False, chunk_batch=1):
        self.model.train()
        return self.forward(data_loader, training=True, average_output=average_output, chunk_batch=chunk_batch)

    def validate(self, data_loader, average_output=False):
        self.model.eval()
        with torch.no_grad():
            return self.forward(data_loader, average_output=average_output, training=False)

    def calibrate_bn(self, data_loader, num_steps=None):
        """Re-estimate BN running statistics as a cumulative average over the loader (momentum=None)."""
        for m in self.model.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.momentum = None
                m.track_running_stats = True
                m.reset_running_stats()
        self.model.train()
        with torch.no_grad():
            return self.forward(data_loader, num_steps=num_steps, training=False)

    # tensorwatch hooks of the reference (trainer.py:287-337) are observability extras, not part of the path
    def set_watcher(self, filename, port=0):
        return False
