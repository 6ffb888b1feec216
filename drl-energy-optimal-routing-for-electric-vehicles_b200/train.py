"""REINFORCE step, baselines and the data-parallel plumbing around the rollout kernels.

Mirrors (same names, argument meaning, return values) the parts of the reference that sit directly on the hot
path:
    rollout                 utils/reinforce_baselines.py:51-69
    ExponentialBaseline     utils/reinforce_baselines.py:149-173
    RolloutBaseline         utils/reinforce_baselines.py:207-276
    WarmupBaseline          utils/reinforce_baselines.py:95-140
    BaselineDataset         utils/reinforce_baselines.py:279-295
    clip_grad_norms         trainer.py:53-70
    reinforce_step          trainer.py:111-127   (one optimisation step)
    eval_dataset / _eval_dataset / eval_widths   trainer.py:382-466 (test-time greedy / sampled evaluation sweep)

Multi-GPU: one process per GPU, instances sharded by contiguous index range, weights
replicated.  Rollouts need no collective.  Training adds exactly two NCCL collectives: one flat all-reduce of
the 874,112 gradient floats per step and the scalar / paired-difference reductions of the baseline statistics.
"""
import math

import numpy as np
import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Dataset


# ------------------------------------------------------------------------------------------------
# sharding helpers
# ------------------------------------------------------------------------------------------------
def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank=None, world_size=None):
    """contiguous instance range [lo, hi) owned by ``rank`` (remainder spread over the first ranks)"""
    if rank is None:
        rank, world_size = world()
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t):
    """in-place sum over ranks.  NCCL reduces device tensors only: a CPU tensor (baseline statistics that went through
    numpy) is staged on this rank's GPU and copied back."""
    r, w = world()
    if w > 1:
        if not t.is_cuda and dist.get_backend() == "nccl":
            d = t.to(torch.device("cuda", torch.cuda.current_device()))
            dist.all_reduce(d, op=dist.ReduceOp.SUM)
            t.copy_(d)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_min_int(v):
    """min over ranks of a Python int (same staging rule)"""
    r, w = world()
    if w == 1:
        return int(v)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([int(v)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())


def global_mean(x):
    """mean over the GLOBAL batch of a per-rank 1-D tensor (loss / baseline means are global, trainer.py:121)"""
    if world()[1] == 1:
        return x.mean()                              # one kernel; what the reference computes (reinforce_baselines.py:160)
    s = torch.stack([x.sum().double(), torch.full((), float(x.numel()), dtype=torch.float64, device=x.device)])
    allreduce_sum_(s)
    return (s[0] / s[1]).to(x.dtype)


def allreduce_gradients(params, global_batch_scale=1.0):
    """ONE flat bucket: sum the gradients of all ranks (NCCL all-reduce over NVLink), scale, scatter back."""
    r, w = world()
    grads = [p.grad for p in params if p.grad is not None]
    if w == 1 or not grads:
        if global_batch_scale != 1.0:
            for g in grads:
                g.mul_(global_batch_scale)
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if global_batch_scale != 1.0:
        flat.mul_(global_batch_scale)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


# ------------------------------------------------------------------------------------------------
# rollout + baselines
# ------------------------------------------------------------------------------------------------
def rollout(actor, dataset, args):
    """greedy eval-mode costs of a whole dataset -> CPU tensor [len(dataset)]"""
    actor.set_decode_type("greedy")
    actor.eval()
    costs = []
    for bat in DataLoader(dataset, batch_size=args.batch_size):
        with torch.no_grad():
            _, _, R = actor(bat)
        costs.append(R.sum(1).cpu())
    return torch.cat(costs, 0)


class Baseline(object):
    def wrap_dataset(self, dataset):
        return dataset

    def unwrap_batch(self, batch):
        return batch, None

    def eval(self, x, c):
        raise NotImplementedError

    def get_learnable_parameters(self):
        return []

    def epoch_callback(self, model, epoch):
        pass

    def state_dict(self):
        return {}

    def load_state_dict(self, state_dict):
        pass


class NoBaseline(Baseline):
    def eval(self, x, c):
        return 0, 0


class ExponentialBaseline(Baseline):
    """v <- beta v + (1-beta) mean(c), mean over the global batch"""

    def __init__(self, beta):
        self.beta, self.v = beta, None

    def eval(self, x, c):
        mean = global_mean(c.detach())
        self.v = mean if self.v is None else self.beta * self.v + (1. - self.beta) * mean
        self.v = self.v.detach()
        return self.v, 0

    def state_dict(self):
        return {'v': self.v}

    def load_state_dict(self, state_dict):
        self.v = state_dict['v']


class BaselineDataset(Dataset):
    def __init__(self, dataset=None, baseline=None):
        assert len(dataset) == len(baseline)
        self.dataset, self.baseline = dataset, baseline

    def __getitem__(self, item):
        return {'data': self.dataset[item], 'baseline': self.baseline[item]}

    def __len__(self):
        return len(self.dataset)


def paired_ttest_one_sided(candidate, baseline):
    """scipy.stats.ttest_rel(candidate, baseline) from GLOBAL sufficient statistics (n, sum d, sum d^2) so that
    each rank only contributes its shard of the validation set.  Returns (t, p_one_sided)."""
    from scipy import stats
    d = (np.asarray(candidate, np.float64) - np.asarray(baseline, np.float64))
    s = torch.tensor([float(d.size), d.sum(), (d * d).sum()], dtype=torch.float64)
    if world()[1] > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        s = allreduce_sum_(s.to(dev)).cpu()
    n, sd, sdd = s.tolist()
    mean = sd / n
    var = (sdd - n * mean * mean) / (n - 1)
    t = mean / math.sqrt(var / n)
    p = stats.t.sf(abs(t), n - 1) * 2
    return t, p / 2


class RolloutBaseline(Baseline):
    def __init__(self, actor, valid_data, args, epoch=0):
        self.dataset, self.args = valid_data, args
        self._update_model(actor, epoch)

    def _update_model(self, actor, epoch):
        self.actor = actor                      # the reference aliases the live actor (no deepcopy), :217-218
        self.bl_vals = rollout(self.actor, self.dataset, self.args).cpu().numpy()
        self.mean = float(global_mean(torch.from_numpy(self.bl_vals)))
        self.epoch = epoch

    def wrap_dataset(self, dataset):
        return BaselineDataset(dataset, rollout(self.actor, dataset, self.args).view(-1, 1))

    def unwrap_batch(self, batch):
        return batch['data'], batch['baseline'].view(-1)

    def eval(self, x, c):
        with torch.no_grad():
            v, _, _ = self.actor(x)
        return v, 0

    def epoch_callback(self, model, epoch):
        cand = rollout(model, self.dataset, self.args).cpu().numpy()
        cand_mean = float(global_mean(torch.from_numpy(cand)))
        if cand_mean - self.mean < 0:
            t, p_val = paired_ttest_one_sided(cand, self.bl_vals)
            assert t < 0, "T-statistic should be negative"
            if p_val < self.args.bl_alpha:
                self._update_model(model, epoch)

    def state_dict(self):
        return {'model': self.actor.state_dict(), 'epoch': self.epoch}

    def load_state_dict(self, state_dict):
        self.actor.load_state_dict(state_dict['model'])
        self._update_model(self.actor, state_dict['epoch'])


class WarmupBaseline(Baseline):
    def __init__(self, baseline, n_epochs=1, warmup_exp_beta=0.8):
        self.baseline = baseline
        assert n_epochs > 0
        self.warmup_baseline = ExponentialBaseline(warmup_exp_beta)
        self.alpha, self.n_epochs = 0, n_epochs

    def wrap_dataset(self, dataset):
        return self.baseline.wrap_dataset(dataset) if self.alpha > 0 else self.warmup_baseline.wrap_dataset(dataset)

    def unwrap_batch(self, batch):
        return self.baseline.unwrap_batch(batch) if self.alpha > 0 else self.warmup_baseline.unwrap_batch(batch)

    def eval(self, x, c):
        if self.alpha == 1:
            return self.baseline.eval(x, c)
        if self.alpha == 0:
            return self.warmup_baseline.eval(x, c)
        v, l = self.baseline.eval(x, c)
        vw, lw = self.warmup_baseline.eval(x, c)
        return self.alpha * v + (1 - self.alpha) * vw, self.alpha * l + (1 - self.alpha * lw)   # :125 as written

    def epoch_callback(self, model, epoch):
        self.baseline.epoch_callback(model, epoch)
        self.alpha = (epoch + 1) / float(self.n_epochs)

    def state_dict(self):
        return self.baseline.state_dict()

    def load_state_dict(self, state_dict):
        self.baseline.load_state_dict(state_dict)


# ------------------------------------------------------------------------------------------------
# fused REINFORCE loss (csrc/evrp_loss.cu)
# ------------------------------------------------------------------------------------------------
class ReinforceLoss(torch.autograd.Function):
    """loss = mean_global((R.sum(1) - baseline).detach() * ll.sum(1))  (trainer.py:116-121) in ONE pair of kernels:
    the forward also produces d loss / d ll, so the backward is a single scale."""

    @staticmethod
    def forward(ctx, ll, R, baseline, inv_global_batch):
        from . import _lib
        L = _lib.lib()
        ll_c, R_c = ll.detach().contiguous(), R.detach().contiguous()
        B, T = ll_c.shape
        dev = ll_c.device
        bl = torch.as_tensor(baseline, dtype=torch.float32, device=dev).detach().reshape(-1).contiguous()
        assert bl.numel() in (1, B)
        reward, adv, scratch = (torch.empty(B, device=dev) for _ in range(3))
        grad_ll = torch.empty(B, T, device=dev)
        loss = torch.empty(1, device=dev)
        # ``inv_global_batch``: a Python float (1 / global batch size), or a CUDA scalar tensor holding the global batch SIZE
        # (multi-rank: the all-reduced count stays on the device, the kernel takes its reciprocal -- no host read-back)
        n_dev = None
        if torch.is_tensor(inv_global_batch):
            n_dev = inv_global_batch.detach().to(dev, torch.float32).reshape(1).contiguous()
            inv_global_batch = 0.0
        rc = L.evrp_reinforce_loss(_lib.ptr(ll_c), _lib.ptr(R_c), _lib.ptr(bl), int(bl.numel() == 1), B, T,
                                   float(inv_global_batch), _lib.ptr(n_dev), _lib.ptr(reward), _lib.ptr(adv), _lib.ptr(grad_ll),
                                   _lib.ptr(loss), _lib.ptr(scratch), _lib.stream_ptr())
        _lib.check(rc, "evrp_reinforce_loss")
        ctx.save_for_backward(grad_ll)
        ctx.mark_non_differentiable(reward, adv)
        return loss[0], reward, adv

    @staticmethod
    def backward(ctx, g_loss, g_reward, g_adv):
        (grad_ll,) = ctx.saved_tensors
        return grad_ll * g_loss, None, None, None


# ------------------------------------------------------------------------------------------------
# one optimisation step
# ------------------------------------------------------------------------------------------------
def clip_grad_norms(param_groups, max_norm=math.inf):
    """trainer.py:53-70.  The clipped norms stay on the device (torch.clamp, not Python's min(): comparing a CUDA tensor
    with a float would stall the host until the whole backward has run)."""
    grad_norms = [torch.nn.utils.clip_grad_norm_(g['params'], max_norm if max_norm > 0 else math.inf, norm_type=2)
                  for g in param_groups]
    clipped = [torch.clamp(g, max=max_norm) for g in grad_norms] if max_norm > 0 else grad_norms
    return grad_norms, clipped


def reinforce_step(actor, baseline, optimizer, batch, max_grad_norm=2.0, forced_actions=None):
    """trainer.py:113-127 for one batch (this rank's shard of the global batch).
    loss = mean over the GLOBAL batch of (reward - baseline).detach() * ll.sum(1)."""
    r, w = world()
    x, bl_val = baseline.unwrap_batch(batch)
    pi, ll, R = actor(x) if forced_actions is None else actor(x, forced_actions=forced_actions)
    reward = R.sum(1)
    if bl_val is None:
        bl_val, bl_loss = baseline.eval(x, reward)
    else:
        bl_val, bl_loss = bl_val.to(reward.device), 0
    n_global = 1.0 / float(reward.numel())
    if w > 1:                                    # ranks may hold different batch sizes: the mean is over the GLOBAL batch;
        n_global = torch.full((), float(reward.numel()), device=reward.device)     # the count stays on the device
        allreduce_sum_(n_global)
    # fused kernel: reward / advantage / loss and d loss / d ll; each rank holds its share of the GLOBAL mean
    loss, reward, advantage = ReinforceLoss.apply(ll, R, bl_val, n_global)
    loss = loss + bl_loss
    optimizer.zero_grad()
    loss.backward()
    allreduce_gradients([p for g in optimizer.param_groups for p in g['params']])
    grad_norms = clip_grad_norms(optimizer.param_groups, max_grad_norm)
    optimizer.step()
    loss_global = allreduce_sum_(loss.detach().clone())
    return dict(loss=loss_global, reward=global_mean(reward.detach()), grad_norm=grad_norms[0][0], pi=pi, T=pi.shape[1])


# ------------------------------------------------------------------------------------------------
# per-rank trainer shell: trainer.py:21-49 validate, :73-206 train, :160-169 checkpoint format
# ------------------------------------------------------------------------------------------------
def rank_shard(dataset):
    """this rank's contiguous slice of a dataset (instances are sharded)"""
    from torch.utils.data import Subset
    lo, hi = shard_bounds(len(dataset))
    return Subset(dataset, range(lo, hi)) if (lo, hi) != (0, len(dataset)) else dataset


def validate(dataset, actor, batch_size):
    """trainer.py:21-49 without the plotting: mean greedy cost of a validation set (mean of per-batch means, as the
    reference computes it; identical to the global mean when the batches are equal-sized).  Sharded over ranks."""
    was_training = actor.training
    actor.eval()
    actor.set_decode_type("greedy")
    sums = torch.zeros(2, dtype=torch.float64)
    for batch in DataLoader(rank_shard(dataset), batch_size, False, num_workers=0):
        with torch.no_grad():
            _, _, R = actor(batch)
        sums += torch.tensor([float(R.sum(1).mean()), 1.0], dtype=torch.float64)
    r, w = world()
    if w > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        sums = allreduce_sum_(sums.to(dev)).cpu()
    actor.train(was_training)
    return float(sums[0] / max(float(sums[1]), 1.0))


def checkpoint_state(actor, optimizer, baseline):
    """the reference's checkpoint dictionary, key for key (trainer.py:160-169)"""
    return {'model': actor.state_dict(), 'optimizer': optimizer.state_dict(), 'rng_state': torch.get_rng_state(),
            'cuda_rng_state': torch.cuda.get_rng_state_all() if torch.cuda.is_available() else [],
            'baseline': baseline.state_dict(),
            # not in the reference's dictionary (it ignores unknown keys): the counter-based sampler of the rollout kernel
            'sampler_state': actor.sampler_state() if hasattr(actor, "sampler_state") else None}


def load_checkpoint(path, actor, optimizer=None, baseline=None, map_location=None):
    """trainer.py:277-301: restore model (+ optimizer, RNG, baseline) from a checkpoint of either code base"""
    ck = torch.load(path, map_location=map_location, weights_only=False)
    actor.load_state_dict({**actor.state_dict(), **ck.get('model', {})})
    if optimizer is not None and 'optimizer' in ck:
        optimizer.load_state_dict(ck['optimizer'])
    if 'rng_state' in ck:
        torch.set_rng_state(ck['rng_state'].cpu() if hasattr(ck['rng_state'], "cpu") else ck['rng_state'])
    if ck.get('cuda_rng_state') and torch.cuda.is_available() and len(ck['cuda_rng_state']) == torch.cuda.device_count():
        torch.cuda.set_rng_state_all([t.cpu() for t in ck['cuda_rng_state']])
    if ck.get('sampler_state') is not None and hasattr(actor, "load_sampler_state"):
        actor.load_sampler_state(ck['sampler_state'])
    if baseline is not None and 'baseline' in ck:
        baseline.load_state_dict(ck['baseline'])
    return ck


def train_epoch(actor, baseline, optimizer, train_data, batch_size, max_grad_norm=2.0, log_every=100, on_log=None,
                shuffle=True):
    """One epoch of trainer.py:100-143 on this rank's shard: wrap the dataset with the baseline values (for the rollout
    baseline: a greedy pass of the rollout kernel over the shard), then REINFORCE steps with a flat gradient all-reduce.
    ``batch_size`` is the PER-RANK batch.  Returns (mean reward, mean loss, steps)."""
    data = baseline.wrap_dataset(rank_shard(train_data))
    loader = DataLoader(data, batch_size, shuffle, num_workers=0)
    # every REINFORCE step issues collectives (global batch size, BatchNorm statistics, gradient all-reduce): all ranks must
    # take the SAME number of steps.  Shards differ by at most one instance, so the batch counts can differ by one: stop at
    # the minimum over ranks (the surplus instances of the longer shards sit out this epoch; shuffling rotates them)
    n_steps = allreduce_min_int(len(loader))
    actor.train()
    actor.set_decode_type("sample")
    rewards, losses = [], []
    for batch_idx, batch in enumerate(loader):
        if batch_idx >= n_steps:
            break
        out = reinforce_step(actor, baseline, optimizer, batch, max_grad_norm)
        rewards.append(float(out["reward"]))
        losses.append(float(out["loss"]))
        if on_log is not None and (batch_idx + 1) % log_every == 0:
            on_log(batch_idx, len(loader), float(np.mean(rewards[-log_every:])), float(np.mean(losses[-log_every:])))
    return float(np.mean(rewards)) if rewards else float("nan"), float(np.mean(losses)) if losses else float("nan"), len(rewards)


def train(actor, baseline, optimizer, lr_scheduler, train_data, valid_data, batch_size, max_grad_norm, iterations,
          save_dir=None, log=print):
    """trainer.py:73-206 as a per-rank loop: epochs of train_epoch -> checkpoint (rank 0, reference format) -> validate
    -> baseline.epoch_callback -> lr_scheduler.step -> best.pt.  CSV rows [reward, loss, seconds, valid] like
    trainer.py:177-179 go to ``save_dir/epochs.csv`` on rank 0."""
    import csv
    import os
    import time
    rank, _ = world()
    best, epoch_reward, epoch_loss = float("inf"), [], []
    if save_dir is not None and rank == 0:
        os.makedirs(os.path.join(save_dir, "checkpoints"), exist_ok=True)
    for epoch in range(iterations):
        t0 = time.time()
        mean_reward, mean_loss, steps = train_epoch(
            actor, baseline, optimizer, train_data, batch_size, max_grad_norm,
            on_log=(lambda i, n, r, l: log('Batch %d/%d, reward: %2.3f,  loss: %2.4f' % (i, n, r, l))) if rank == 0 else None)
        epoch_reward.append(mean_reward); epoch_loss.append(mean_loss)
        if save_dir is not None and rank == 0:
            d = os.path.join(save_dir, "checkpoints", str(epoch))
            os.makedirs(d, exist_ok=True)
            torch.save(checkpoint_state(actor, optimizer, baseline), os.path.join(d, 'epoch-{}.pt'.format(epoch)))
        mean_valid = validate(valid_data, actor, batch_size)
        if save_dir is not None and rank == 0:
            with open(os.path.join(save_dir, "epochs.csv"), "a", newline='') as f:
                csv.writer(f).writerow([mean_reward, mean_loss, time.time() - t0, mean_valid])
        baseline.epoch_callback(actor, epoch)
        if lr_scheduler is not None:
            lr_scheduler.step()
        if mean_valid < best:
            best = mean_valid
            if save_dir is not None and rank == 0:
                torch.save(checkpoint_state(actor, optimizer, baseline), os.path.join(save_dir, 'best.pt'))
        if rank == 0:
            log('Mean epoch loss/reward: %2.4f, %2.4f, %2.4f, took: %2.4fs (%d steps)' %
                (mean_loss, mean_reward, mean_valid, time.time() - t0, steps))
    return epoch_reward, epoch_loss


# ------------------------------------------------------------------------------------------------
# test-time evaluation: trainer.py:382-384 width sweep, :388-419 eval_dataset, :421-466 _eval_dataset
# ------------------------------------------------------------------------------------------------
def _eval_dataset(model, dataset, width, softmax_temp, args):
    """trainer.py:421-466: per batch ``sample_many`` with (batch_rep, iter_rep) derived from ``width`` exactly as the
    reference does (greedy: width 0, one replica; sample: ``width`` replicas, split into ``iter_rep`` passes of
    ``max_calc_batch_size`` when they do not fit).  -> [(costs [b] (device), seconds)].  Like the reference,
    ``softmax_temp`` is accepted but NOT applied (``set_decode_type`` is called without it, trainer.py:424-425);
    pass ``args.apply_softmax_temperature = True`` to apply it.  Beam search ('bs') is unreachable in the reference
    (``self.problem`` undefined, nets/DRLModel.py:143) and raises here."""
    import time
    model.eval()
    greedy = args.decode_strategy in ('bs', 'greedy')
    if getattr(args, "apply_softmax_temperature", False):
        model.set_decode_type("greedy" if greedy else "sample", temp=softmax_temp)
    else:
        model.set_decode_type("greedy" if greedy else "sample")
    results = []
    for batch in DataLoader(dataset, args.eval_batch_size, False, num_workers=0):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        start = time.time()
        with torch.no_grad():
            if args.decode_strategy == 'greedy':
                assert width == 0, "Do not set width when using greedy"
                assert args.eval_batch_size <= args.max_calc_batch_size, \
                    "eval_batch_size should be smaller than calc batch size"
                batch_rep, iter_rep = 1, 1
            elif args.decode_strategy == 'sample':
                if width * args.eval_batch_size > args.max_calc_batch_size:
                    assert args.eval_batch_size == 1
                    assert width % args.max_calc_batch_size == 0
                    batch_rep, iter_rep = args.max_calc_batch_size, width // args.max_calc_batch_size
                else:
                    batch_rep, iter_rep = width, 1
                assert batch_rep > 0
            else:
                raise NotImplementedError("decode_strategy 'bs': beam search is broken in the reference itself")
            costs, _ = model.sample_many(batch, batch_rep=batch_rep, iter_rep=iter_rep)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        results.append((costs, time.time() - start))
    return results


def eval_dataset(test_data, width, softmax_temp, args, actor, csv_path=None):
    """trainer.py:388-419 -> (mean cost, durations).  ``csv_path`` (optional) receives the reference's rows
    [cost, batch duration] + the mean row; the reference's hard-coded ExperimentalLog/ path is not reproduced."""
    import csv
    results = _eval_dataset(actor, test_data, width, softmax_temp, args)
    costs, durations = zip(*results)
    costs = torch.cat(costs, 0).cpu().numpy()
    if csv_path is not None:
        with open(csv_path, "a", newline='') as f:
            w = csv.writer(f)
            per = len(costs) // max(len(durations), 1)
            for i in range(len(costs)):
                w.writerow([costs[i], durations[min(i // max(per, 1), len(durations) - 1)]])
            w.writerow(["mean", float(np.mean(costs)), float(np.mean(durations))])
    return float(np.mean(costs)), durations


def eval_widths(test_data, args, actor, csv_path=None):
    """trainer.py:382-384: one evaluation per entry of ``args.width`` (None -> [0])"""
    widths = args.width if getattr(args, "width", None) is not None else [0]
    return [eval_dataset(test_data, w, getattr(args, "softmax_temperature", 1.0), args, actor, csv_path)[0] for w in widths]
