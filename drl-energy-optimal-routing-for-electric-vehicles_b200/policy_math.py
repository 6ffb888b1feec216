"""Host-side helpers of the drop-in model:

  pack_weights        reference state_dict layout -> kernel layouts (k-major, folded BN / project_out / tour)
  pairwise_matrices   the 3-tuple input branch of nets/DRLModel.py:124-133 (distance / clamped-slope build)
  encode_torch        differentiable encoder + precompute (training forward; train-mode BatchNorm)
  teacher_forced_ll   differentiable log-likelihood of a recorded rollout, all T steps batched at once

The two differentiable functions are plain PyTorch on the GPU; they exist so that autograd reaches all 50
parameters for the REINFORCE step (trainer.py:115-127) while the rollout itself runs in the CUDA kernel.
They are NOT used by the greedy / sampled inference path.
"""
import math

import torch
import torch.nn.functional as Fn

from . import _lib


def _bn_fold(bn):
    """eval-mode BatchNorm1d as y = x*scale + shift (nets/GraphEncoder.py:144-145), folded in fp64"""
    w, b = bn.weight.double(), bn.bias.double()
    rm, rv = bn.running_mean.double(), bn.running_var.double()
    scale = w / torch.sqrt(rv + bn.eps)
    shift = b - rm * scale
    return scale.float().contiguous(), shift.float().contiguous()


def folded_node_projection(model):
    """[384,128]: rows 0..255 = glimpse K|V projection, rows 256..383 = project_out^T @ W_logitK so that
    logits[j] = heads . lK[j]  (nets/DRLModel.py:436-444).  Differentiable."""
    W = model.project_node_embeddings.weight
    Wl = (model.project_out.weight.double().t() @ W[256:].double()).to(W.dtype)
    return torch.cat([W[:256], Wl], 0)


def folded_tour_weight(model):
    """[512,3] = FF_tour.0.weight[:, :128] @ tour.weight (nets/DRLModel.py:233-235).  Differentiable."""
    W1 = model.FF_tour[0].weight
    return (W1[:, :128].double() @ model.tour.weight.double()).to(W1.dtype)


def split_tf32(w):
    """w (fp32) -> (hi, lo) TF32 planes with round-to-nearest (ties away, = PTX cvt.rna.tf32.f32): w = hi + lo up to
    2^-24 |w|.  Both planes have their 13 low mantissa bits cleared, so the tensor core reads them exactly."""
    def rna(x):
        bits = x.contiguous().view(torch.int32)
        return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    w = w.detach().float().contiguous()
    hi = rna(w)
    lo = rna(w - hi)
    return hi.contiguous(), lo.contiguous()


def split_f16(w):
    """w (fp32) -> (hi, lo, inv_scale): fp16 planes of s*w with s a power of two chosen so that max|s*w| ~ 2^12
    (keeps the lo plane out of the fp16 subnormal range); s*w = hi + lo to 22 significant bits."""
    w = w.detach().float().contiguous()
    amax = float(w.abs().max()) if w.numel() else 1.0
    e = 12 - math.frexp(amax)[1] if amax > 0 else 0
    e = max(-14, min(24, e))
    ws = w * (2.0 ** e)                               # exact
    hi = ws.half()
    lo = (ws - hi.float()).half()
    return hi.contiguous(), lo.contiguous(), 2.0 ** (-e)


def pack_ff_stream(w1_hi, w1_lo, w2_hi, w2_lo):
    """The four fp16 planes of FF_tour (W1[:, 128:] [512,128], W2 [128,512]) as the ONE 512 KB stream the rollout kernel
    consumes (include/evrp_b200.h ``ff_stream``): 16 stages x (hi | lo) of [128 rows][64 halfs] tiles with the 128-byte
    shared-memory swizzle applied (16-byte chunk c of row r stored at c ^ (r & 7))."""
    def tiles(p, n_row_tiles, n_kb):          # [rows, K] -> [row tile, K block, 128, 8 chunks, 8]
        t = p.reshape(n_row_tiles, 128, n_kb, 8, 8).permute(0, 2, 1, 3, 4)
        r = torch.arange(128, device=p.device)[:, None]
        c = torch.arange(8, device=p.device)[None, :]
        src = (c ^ (r & 7))[None, None, :, :, None].expand_as(t)          # dest chunk c holds source chunk c ^ (r & 7)
        return torch.gather(t, 3, src)
    a = torch.stack([tiles(w1_hi, 4, 2), tiles(w1_lo, 4, 2)], 2)           # [4, 2, plane, 128, 8, 8]
    b = torch.stack([tiles(w2_hi, 1, 8), tiles(w2_lo, 1, 8)], 2)           # [1, 8, plane, 128, 8, 8]
    return torch.cat([a.reshape(-1), b.reshape(-1)]).contiguous()


FUSED_LP = 1280          # floats of fused_params per layer (csrc/evrp_encoder_fused.cu)


def _fused_blocks(w, chunks, kblocks):
    """w [N,K] fp32 (nn.Linear layout) -> (list of fp16 [128,64] blocks, 1/scale).  Order: for every (chunk of 128
    output rows, K block of 64 inputs) in ``order``: hi block then lo block."""
    hi, lo, inv = split_f16(w)
    out = []
    for c, kb in zip(chunks, kblocks):
        out.append(hi[128 * c:128 * c + 128, 64 * kb:64 * kb + 64])
        out.append(lo[128 * c:128 * c + 128, 64 * kb:64 * kb + 64])
    return out, inv


def pack_fused_encoder(model, layer_mats, node_mat, bn_folds):
    """Operand stream of the fused encoder kernel (include/evrp_b200.h ``fused_w`` / ``fused_params``).
    layer_mats: per layer (Wqkv [384,128], Wo [128,128], W1 [512,128], b1, W2 [128,512], b2) in nn.Linear layout;
    node_mat [384 or 512, 128]; bn_folds: per layer ((scale1, shift1), (scale2, shift2))."""
    blocks, params = [], []
    for (Wqkv, Wo, W1, b1, W2, b2), ((s1, h1), (s2, h2)) in zip(layer_mats, bn_folds):
        bq, iq = _fused_blocks(Wqkv, [0, 0, 1, 1, 2, 2], [0, 1, 0, 1, 0, 1])
        bo, io = _fused_blocks(Wo, [0, 0], [0, 1])
        w1h, w1l, i1 = split_f16(W1)
        w2h, w2l, i2 = split_f16(W2)

        def ff1(c):
            return [t[128 * c:128 * c + 128, 64 * kb:64 * kb + 64] for kb in (0, 1) for t in (w1h, w1l)]

        def ff2(j):
            return [t[:, 64 * kb:64 * kb + 64] for kb in (2 * j, 2 * j + 1) for t in (w2h, w2l)]
        blocks += bq + bo + ff1(0) + ff1(1) + ff2(0) + ff1(2) + ff2(1) + ff1(3) + ff2(2) + ff2(3)
        p = torch.zeros(FUSED_LP, dtype=torch.float32, device=Wqkv.device)
        p[0:128], p[128:256], p[256:768], p[768:896], p[896:1024], p[1024:1152] = s1, h1, b1, b2, s2, h2
        p[1152:1156] = torch.tensor([iq, io, i1, i2], dtype=torch.float32)
        params.append(p)
    n_chunks = node_mat.shape[0] // 128
    bn, inode = _fused_blocks(node_mat, [c for c in range(n_chunks) for _ in (0, 1)], [0, 1] * n_chunks)
    blocks += bn
    params.append(torch.full((1,), inode, dtype=torch.float32, device=node_mat.device))
    W = torch.stack([b.contiguous() for b in blocks]).contiguous()         # [n_blocks,128,64] fp16
    return W, torch.cat(params).contiguous(), n_chunks


def pack_weights(model):
    """both halves (encoder + decoder) in one dictionary"""
    keep = pack_encoder_weights(model)
    keep.update(pack_decoder_weights(model))
    return keep


def pack_encoder_weights(model):
    """kernel-side layouts of the encoder / precompute weights -> {"enc_struct": EncoderWeights, tensors kept alive}"""
    keep = {}
    ew = _lib.EncoderWeights()

    def dev(name, t):
        t = t.detach().float().contiguous()
        keep[name] = t
        return _lib.ptr(t) if t.is_cuda else None

    ew.depot_w = dev("depot_w", model.init_embed_depot_and_station.weight)
    ew.depot_b = dev("depot_b", model.init_embed_depot_and_station.bias)
    ew.station_w = dev("station_w", model.init_embed_station.weight)
    ew.station_b = dev("station_b", model.init_embed_station.bias)
    ew.cust_w = dev("cust_w", model.init_embed.weight)
    ew.cust_b = dev("cust_b", model.init_embed.bias)
    layers = model.embedder.layers
    ew.n_layers = len(layers)
    gemm_mode = int(getattr(model, "gemm_mode", 0))
    fused_layers, fused_bn = [], []
    for l, layer in enumerate(layers):
        att, bn1, ff, bn2 = layer[0].module, layer[1].normalizer, layer[2].module, layer[3].normalizer
        Hh, d, dk = att.W_query.shape
        cols = [w.permute(1, 0, 2).reshape(d, Hh * dk) for w in (att.W_query, att.W_key, att.W_val)]
        lw = ew.layers[l]
        lw.Wqkv = dev(f"l{l}.Wqkv", torch.cat(cols, 1))
        lw.Wo = dev(f"l{l}.Wo", att.W_out.reshape(Hh * dk, d))
        s, sh = _bn_fold(bn1)
        lw.bn1_scale, lw.bn1_shift = dev(f"l{l}.bn1s", s), dev(f"l{l}.bn1b", sh)
        lw.ff1_w, lw.ff1_b = dev(f"l{l}.ff1w", ff[0].weight.t()), dev(f"l{l}.ff1b", ff[0].bias)
        lw.ff2_w, lw.ff2_b = dev(f"l{l}.ff2w", ff[2].weight.t()), dev(f"l{l}.ff2b", ff[2].bias)
        s, sh = _bn_fold(bn2)
        lw.bn2_scale, lw.bn2_shift = dev(f"l{l}.bn2s", s), dev(f"l{l}.bn2b", sh)
        fused_layers.append((torch.cat(cols, 1).t().float(), att.W_out.reshape(Hh * dk, d).t().float(), ff[0].weight.float(),
                             ff[0].bias.float(), ff[2].weight.float(), ff[2].bias.float()))
        fused_bn.append((_bn_fold(bn1), (s, sh)))
        # tcgen05 path: [out][in] planes
        for name, mat in (("Wqkv", torch.cat(cols, 1).t()), ("Wo", att.W_out.reshape(Hh * dk, d).t()),
                          ("ff1", ff[0].weight), ("ff2", ff[2].weight)):
            hi, lo = split_tf32(mat)
            setattr(lw, name + "_hi", dev(f"l{l}.{name}_hi", hi))
            setattr(lw, name + "_lo", dev(f"l{l}.{name}_lo", lo))
    hi, lo = split_tf32(folded_node_projection(model))
    ew.node_proj_hi, ew.node_proj_lo = dev("node_proj_hi", hi), dev("node_proj_lo", lo)
    ew.gemm_mode = gemm_mode
    ew.attn_mode = int(getattr(model, "attn_mode", 0))
    ew.node_proj = dev("node_proj", folded_node_projection(model).t())
    ew.fixed_ctx_w = dev("fixed_ctx_w", model.project_fixed_context.weight.t())
    policy = int(getattr(model, "policy", 0))
    node_mat = folded_node_projection(model).float()
    if policy == 1:                                    # the AM step table rides behind the node projection (4th chunk)
        node_mat = torch.cat([node_mat, model.project_step_context.weight[:, :128].float()], 0)
    fw, fp, nchunks = pack_fused_encoder(model, fused_layers, node_mat, fused_bn)
    keep["fused_w"], keep["fused_params"] = fw, fp
    ew.fused_w = _lib.ptr(fw) if fw.is_cuda else None
    ew.fused_params = _lib.ptr(fp) if fp.is_cuda else None
    ew.fused_node_chunks = nchunks
    if policy == 1:                                    # nets/AM.py:90: project_step_context = [emb columns | load column]
        Ws = model.project_step_context.weight
        ew.step_proj = dev("step_proj", Ws[:, :128].t())
        hi, lo = split_tf32(Ws[:, :128])
        ew.step_proj_hi, ew.step_proj_lo = dev("step_proj_hi", hi), dev("step_proj_lo", lo)
    keep["enc_struct"] = ew
    return keep


def pack_decoder_weights(model):
    """kernel-side layouts of the decoder (FF_tour / step-context) weights -> {"dec_struct": DecoderWeights, tensors}.
    The REINFORCE step needs only this half every optimizer step (its encoder runs through encode_torch)."""
    keep = {}

    def dev(name, t):
        t = t.detach().float().contiguous()
        keep[name] = t
        return _lib.ptr(t) if t.is_cuda else None

    dw = _lib.DecoderWeights()
    dw.policy = int(getattr(model, "policy", 0))
    if dw.policy == 1:
        dw.am_w_load = dev("am_w_load", model.project_step_context.weight[:, 128])
        dw.ff_mode = 1
        keep["dec_struct"] = dw
        return keep
    W1, W2, tw = model.FF_tour[0].weight, model.FF_tour[2].weight, model.tour.weight
    if W1.is_cuda and W1.dtype == torch.float32 and tuple(W1.shape) == (512, 256) and tuple(W2.shape) == (128, 512) \
            and tuple(tw.shape) == (128, 3) and W1.is_contiguous() and W2.is_contiguous() and tw.is_contiguous():
        # product path: two kernels, no host read-back (csrc/evrp_train.cu::pack_decoder_kernel); same values as below
        d = W1.device
        bufs = dict(w1_veh=torch.empty(3, 512, device=d), w1_max=torch.empty(128, 512, device=d), w2=torch.empty(512, 128, device=d),
                    w1max_hi=torch.empty(512, 128, device=d, dtype=torch.float16), w1max_lo=torch.empty(512, 128, device=d, dtype=torch.float16),
                    w2_hi=torch.empty(128, 512, device=d, dtype=torch.float16), w2_lo=torch.empty(128, 512, device=d, dtype=torch.float16),
                    inv_scales=torch.empty(2, device=d), pack_scratch=torch.empty(2, device=d),
                    ff_stream=torch.empty(16 * 16384, device=d, dtype=torch.float16))
        keep.update(bufs)
        _lib.check(_lib.lib().evrp_pack_decoder(
            _lib.ptr(tw.detach()), _lib.ptr(W1.detach()), _lib.ptr(W2.detach()), _lib.ptr(bufs["w1_veh"]), _lib.ptr(bufs["w1_max"]),
            _lib.ptr(bufs["w2"]), _lib.ptr(bufs["w1max_hi"]), _lib.ptr(bufs["w1max_lo"]), _lib.ptr(bufs["w2_hi"]), _lib.ptr(bufs["w2_lo"]),
            _lib.ptr(bufs["inv_scales"]), _lib.ptr(bufs["pack_scratch"]), _lib.ptr(bufs["ff_stream"]), _lib.stream_ptr()),
            "evrp_pack_decoder")
        dw.ff_stream = _lib.ptr(bufs["ff_stream"])
        dw.w1_veh, dw.w1_max, dw.w2 = _lib.ptr(bufs["w1_veh"]), _lib.ptr(bufs["w1_max"]), _lib.ptr(bufs["w2"])
        dw.b1, dw.b2 = dev("b1", model.FF_tour[0].bias), dev("b2", model.FF_tour[2].bias)
        dw.w1max_hi, dw.w1max_lo = _lib.ptr(bufs["w1max_hi"]), _lib.ptr(bufs["w1max_lo"])
        dw.w2_hi, dw.w2_lo = _lib.ptr(bufs["w2_hi"]), _lib.ptr(bufs["w2_lo"])
        dw.w1_inv_scale = dw.w2_inv_scale = 1.0
        dw.inv_scales = _lib.ptr(bufs["inv_scales"])
        dw.ff_mode = int(getattr(model, "ff_mode", 0))
        keep["dec_struct"] = dw
        return keep
    dw.w1_veh = dev("w1_veh", folded_tour_weight(model).t())
    dw.w1_max = dev("w1_max", model.FF_tour[0].weight[:, 128:].t())
    dw.b1 = dev("b1", model.FF_tour[0].bias)
    dw.w2 = dev("w2", model.FF_tour[2].weight.t())
    dw.b2 = dev("b2", model.FF_tour[2].bias)

    def dev_raw(name, t):                              # fp16 planes travel as they are
        keep[name] = t
        return _lib.ptr(t) if t.is_cuda else None
    hi, lo, inv = split_f16(model.FF_tour[0].weight[:, 128:])
    dw.w1max_hi, dw.w1max_lo, dw.w1_inv_scale = dev_raw("w1max_hi", hi), dev_raw("w1max_lo", lo), inv
    hi2, lo2, inv = split_f16(model.FF_tour[2].weight)
    dw.w2_hi, dw.w2_lo, dw.w2_inv_scale = dev_raw("w2_hi", hi2), dev_raw("w2_lo", lo2), inv
    if tuple(hi.shape) == (512, 128) and tuple(hi2.shape) == (128, 512):
        dw.ff_stream = dev_raw("ff_stream", pack_ff_stream(hi, lo, hi2, lo2))
    dw.ff_mode = int(getattr(model, "ff_mode", 0))
    keep["dec_struct"] = dw
    return keep


def pairwise_matrices(static, elevations):
    """nets/DRLModel.py:128-133 / problems/EVRP.py:87-92.  CUDA tensors: one fused kernel (evrp_pairwise_build);
    CPU tensors (host-side dataset construction): vectorised torch, no Python loop over S."""
    if static.is_cuda:
        static = static.float().contiguous()
        elevations = elevations.to(static.device, torch.float32).contiguous()
        B, _, S = static.shape
        D = torch.empty(B, S, S, device=static.device)
        G = torch.empty(B, S, S, device=static.device)
        _lib.check(_lib.lib().evrp_pairwise_build(_lib.ptr(static), _lib.ptr(elevations), B, S, _lib.ptr(D), _lib.ptr(G),
                                                  _lib.stream_ptr()), "evrp_pairwise_build")
        return D, G
    diff = static[:, :, :, None] - static[:, :, None, :]
    D = torch.sqrt((diff * diff).sum(1))
    G = torch.clamp((elevations[:, :, None] - elevations[:, None, :]) / (D + 0.000001), min=-0.10, max=0.10)
    return D.contiguous(), G.contiguous()


def init_embed_torch(model, static, dynamic):
    F = model.charging_num
    xy = static.transpose(1, 2) / 100
    dem = dynamic[:, 1, :, None]
    return torch.cat((model.init_embed_depot_and_station(xy[:, 0:1]),
                      model.init_embed_station(xy[:, 1:F + 1]),
                      model.init_embed(torch.cat((xy[:, F + 1:], dem[:, F + 1:]), 2))), 1)


class InitEmbed(torch.autograd.Function):
    """nets/DRLModel.py:192-200 on the init-embedding kernel (csrc/evrp_encoder.cu) and its backward: the six parameter
    gradients as weighted column reductions of grad_h (K = 2 / 2 / 3: no GEMM)."""

    @staticmethod
    def forward(ctx, static, dynamic, wd, bd, ws, bs, wc, bc, F):
        static, dynamic = static.contiguous(), dynamic.contiguous()
        B, _, S = static.shape
        h = torch.empty(B, S, 128, device=static.device, dtype=torch.float32)
        p = [t.detach().contiguous() for t in (wd, bd, ws, bs, wc, bc)]
        _lib.check(_lib.lib().evrp_init_embed_forward(_lib.ptr(static), _lib.ptr(dynamic), B, S, F, *[_lib.ptr(t) for t in p],
                                                      _lib.ptr(h), _lib.stream_ptr()), "evrp_init_embed_forward")
        ctx.save_for_backward(static, dynamic)
        ctx.F = F
        return h

    @staticmethod
    def backward(ctx, g):
        static, dynamic = ctx.saved_tensors
        B, _, S = static.shape
        L = _lib.lib()
        dev = static.device
        out = [torch.empty(128, 2, device=dev), torch.empty(128, device=dev), torch.empty(128, 2, device=dev),
               torch.empty(128, device=dev), torch.empty(128, 3, device=dev), torch.empty(128, device=dev)]
        ws = torch.empty(L.evrp_init_embed_backward_workspace_bytes(), dtype=torch.uint8, device=dev)
        _lib.check(L.evrp_init_embed_backward(_lib.ptr(static), _lib.ptr(dynamic), _lib.ptr(g.contiguous()), B, S, ctx.F,
                                              *[_lib.ptr(t) for t in out], _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                   "evrp_init_embed_backward")
        return (None, None) + tuple(out) + (None,)


class _GlobalBatchNorm(torch.autograd.Function):
    """Train-mode BatchNorm1d over the GLOBAL batch when instances are sharded over ranks: the batch statistics of
    nets/GraphEncoder.py:144-145 couple every instance of the batch, so each rank contributes [sum x, sum x^2, n]
    (and, backward, [sum g, sum g*xhat]) through one all-reduce of 2*C+1 floats.  Same outputs / gradients /
    running-statistics update (momentum, unbiased variance) as nn.BatchNorm1d on the concatenated batch."""

    @staticmethod
    def forward(ctx, x, weight, bias, bn):
        import torch.distributed as dist
        C = x.shape[1]
        st = torch.cat([x.sum(0), (x * x).sum(0), x.new_tensor([float(x.shape[0])])])
        dist.all_reduce(st, op=dist.ReduceOp.SUM)
        n = st[2 * C]
        mean = st[:C] / n
        var = st[C:2 * C] / n - mean * mean                      # biased (normalisation)
        with torch.no_grad():
            m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
            bn.running_mean.mul_(1 - m).add_(m * mean)
            bn.running_var.mul_(1 - m).add_(m * var * (n / (n - 1)))   # unbiased, like F.batch_norm
            bn.num_batches_tracked += 1
        inv = torch.rsqrt(var + bn.eps)
        xhat = (x - mean) * inv
        ctx.save_for_backward(xhat, weight, inv, n)
        return xhat * weight + bias

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist
        xhat, weight, inv, n = ctx.saved_tensors
        C = g.shape[1]
        st = torch.cat([g.sum(0), (g * xhat).sum(0)])
        gw, gb = st[C:].clone(), st[:C].clone()                  # this rank's share; summed with the other gradients
        dist.all_reduce(st, op=dist.ReduceOp.SUM)
        gx = (weight * inv) * (g - st[:C] / n - xhat * (st[C:] / n))
        return gx, gw, gb, None


_LOCAL_BN = [False]


class local_batch_norm:
    """context: BatchNorm statistics of THIS rank's rows only even inside a process group (bench.py's single-rank
    recompute of the multi-rank gradients)"""

    def __enter__(self):
        _LOCAL_BN[0] = True

    def __exit__(self, *a):
        _LOCAL_BN[0] = False


def _batch_norm(bn, x, training, force_global=False):
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and (dist.get_world_size() > 1 or force_global) and not _LOCAL_BN[0]
    if training and x.is_cuda and x.shape[1] == 128 and x.dtype == torch.float32:
        # product path: the two-phase BatchNorm kernels; one all-reduce of the fp64 sums per pass when sharded over ranks
        return BatchNormTrain.apply(x, bn.weight, bn.bias, bn, multi)
    if training and multi:
        return _GlobalBatchNorm.apply(x, bn.weight, bn.bias, bn)
    return bn(x)


def encode_torch(model, static, dynamic):
    """Differentiable restatement of the encoder kernels' math (nets/GraphEncoder.py:55-118,142-179).
    BatchNorm follows ``model.training`` (batch statistics + running-stat update in train mode; with instances
    sharded over ranks the statistics are those of the global batch, see _GlobalBatchNorm)."""
    if static.is_cuda and model.init_embed.weight.shape[0] == 128:
        h = InitEmbed.apply(static.float(), dynamic.float(), model.init_embed_depot_and_station.weight,
                            model.init_embed_depot_and_station.bias, model.init_embed_station.weight, model.init_embed_station.bias,
                            model.init_embed.weight, model.init_embed.bias, int(model.charging_num))
    else:
        h = init_embed_torch(model, static, dynamic)
    B, S, d = h.shape
    for layer in model.embedder.layers:
        att, bn1, ff, bn2 = layer[0].module, layer[1].normalizer, layer[2].module, layer[3].normalizer
        flat = h.reshape(B * S, d)
        Hh, _, dk = att.W_query.shape
        # one [128,384] projection instead of 3 x 8 narrow ones (same products; autograd scatters the gradient back)
        Wqkv = torch.cat([w.permute(1, 0, 2).reshape(d, Hh * dk) for w in (att.W_query, att.W_key, att.W_val)], 1)
        if flat.is_cuda and Hh == 8 and dk == 16 and d == 128:
            # product path: projections / feed-forward on the tcgen05 3xTF32 GEMM (forward and input gradients), the
            # attention core on the flash-style kernels; BatchNorm (batch statistics) and weight gradients through torch
            qkv = TCLinear.apply(flat, Wqkv.t(), None, None, False)                      # [B*S, Q|K|V x (h, dk)]
            heads_flat = SelfAttention.apply(qkv.view(B, S, 3 * d)).reshape(B * S, d)
            x1 = TCLinear.apply(heads_flat, att.W_out.reshape(-1, d).t(), None, flat, False)        # h + heads @ W_out
            h1 = _batch_norm(bn1, x1, model.training)
            x2 = FeedForwardTC.apply(h1, ff[0].weight, ff[0].bias, ff[2].weight, ff[2].bias, h1)     # h1 + FF(h1)
            h = _batch_norm(bn2, x2, model.training).view(B, S, d)
            continue
        # CPU tensors (tests): the plain PyTorch statement of the same layer
        qkv = torch.mm(flat, Wqkv)
        heads_flat = self_attention_torch(qkv.view(B, S, 3 * d), Hh).reshape(B * S, d)
        out = torch.mm(heads_flat, att.W_out.reshape(-1, d)).view(B, S, d)
        h = _batch_norm(bn1, (h + out).view(-1, d), model.training).view(B, S, d)
        h = _batch_norm(bn2, (h + ff(h)).view(-1, d), model.training).view(B, S, d)
    if h.is_cuda and d == 128:
        kvl = TCLinear.apply(h.reshape(B * S, d), folded_node_projection(model), None, None, False).view(B, S, 3 * d)
    else:
        kvl = Fn.linear(h, folded_node_projection(model))
    if h.is_cuda and d == 128:
        fixed = TCLinear.apply(h.mean(1), model.project_fixed_context.weight, None, None, False)
    else:
        fixed = model.project_fixed_context(h.mean(1))
    return h, kvl, fixed


_AFFINE_ID = {}


def _affine_identity(dev, n):
    key = (str(dev), n)
    if key not in _AFFINE_ID:
        _AFFINE_ID[key] = (torch.ones(n, device=dev), torch.zeros(n, device=dev))
    return _AFFINE_ID[key]


def split_tf32_planes(W):
    """CUDA weight [rows, cols] with ANY strides (a transposed view for the input-gradient product, a column slice) ->
    dense TF32 (hi, lo) planes in ONE kernel (evrp_split_tf32; same rounding as split_tf32)."""
    W = W.detach()
    assert W.is_cuda and W.dtype == torch.float32 and W.dim() == 2
    rows, cols = W.shape
    hi = torch.empty(rows, cols, device=W.device, dtype=torch.float32)
    lo = torch.empty_like(hi)
    _lib.check(_lib.lib().evrp_split_tf32(_lib.C.c_void_p(W.data_ptr()), rows, cols, W.stride(0), W.stride(1), _lib.ptr(hi),
                                          _lib.ptr(lo), _lib.stream_ptr()), "evrp_split_tf32")
    return hi, lo


def tc_gemm(x, W, bias=None, resid=None, relu=False, aux=None, aux_w=None, gate=None):
    """((x @ W^T + bias + aux @ aux_w^T) [relu] + resid) on the tcgen05 3xTF32 GEMM of csrc/evrp_tc_gemm.cu
    (evrp_gemm_3xtf32).  x [M,K], W [N,K] (nn.Linear layout, any strides), aux [M,3], aux_w [N,3], fp32 CUDA.
    Epilogues the kernel does not instantiate are finished in torch."""
    x = x.contiguous()
    M, K = x.shape
    N = W.shape[0]
    assert W.shape[1] == K and N % 128 == 0 and K % 32 == 0, (tuple(x.shape), tuple(W.shape))
    hi, lo = split_tf32_planes(W)
    y = torch.empty(M, N, device=x.device, dtype=torch.float32)
    one, zero = _affine_identity(x.device, N)
    epi, post_bias, post_relu, post_resid = 0, None, False, None
    if gate is not None:                                   # y * (gate > 0): input gradient through a ReLU (gate = its output)
        assert bias is None and resid is None and not relu and aux is None and K == 128 and gate.shape == (M, N)
        epi, resid = 32, gate
    elif aux is not None:
        assert bias is not None and relu and resid is None and K == 128 and aux.shape == (M, 3) and aux_w.shape == (N, 3)
        epi = 1 | 2 | 16
    elif bias is not None and relu and resid is None and K == 128:
        epi = 1 | 2
    elif bias is None and not relu and resid is not None and K == 128:
        epi = 4 | 8
    elif bias is not None and not relu and resid is not None and K == 512:
        epi = 1 | 4 | 8
    else:
        post_bias, post_relu, post_resid = bias, relu, resid
    b = bias.detach().contiguous() if (epi & 1) else None
    r = resid.detach().contiguous() if (epi & (4 | 32)) else None
    a = aux.detach().contiguous() if (epi & 16) else None
    aw = aux_w.detach().t().contiguous() if (epi & 16) else None          # [3][N]
    _lib.check(_lib.lib().evrp_gemm_3xtf32(epi, _lib.ptr(x), K, _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(y), N, M, N, K,
                                           _lib.ptr(b), _lib.ptr(r), N, _lib.ptr(one) if epi & 8 else None,
                                           _lib.ptr(zero) if epi & 8 else None, _lib.ptr(a), _lib.ptr(aw),
                                           _lib.stream_ptr()), "evrp_gemm_3xtf32")
    if post_bias is not None:
        y = y + post_bias
    if post_relu:
        y = torch.relu(y)
    if post_resid is not None:
        y = y + post_resid
    return y


_WGRAD_WS = {}


def tc_wgrad(gy, x, want_bias=False, aux=None):
    """gW [N,K] = gy[M,N]^T @ x[M,K] on the tcgen05 split-K weight-gradient kernel (csrc/evrp_wgrad.cu,
    evrp_wgrad_3xtf32; fp32-faithful 3xTF32, deterministic reduction).  Also returns gaux [(1 + n_aux), N]: row 0 = the bias
    gradient gy.sum(0), rows 1.. = aux^T gy for aux [M, n_aux <= 3] (None when neither is wanted)."""
    gy, x = gy.contiguous(), x.contiguous()
    M, N = gy.shape
    K = x.shape[1]
    L = _lib.lib()
    key = str(gy.device)
    ws = _WGRAD_WS.get(key)
    if ws is None:
        ws = _WGRAD_WS[key] = torch.empty(L.evrp_wgrad_workspace_bytes(), dtype=torch.uint8, device=gy.device)
    n_aux = 0 if aux is None else aux.shape[1]
    gW = torch.empty(N, K, device=gy.device, dtype=torch.float32)
    gaux = torch.empty(1 + n_aux, N, device=gy.device, dtype=torch.float32) if (want_bias or n_aux) else None
    a = aux.detach().contiguous() if n_aux else None
    _lib.check(L.evrp_wgrad_3xtf32(_lib.ptr(gy), N, N, _lib.ptr(x), K, K, M, _lib.ptr(a), n_aux, _lib.ptr(gW), _lib.ptr(gaux), 0,
                                   _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "evrp_wgrad_3xtf32")
    return gW, gaux


class TCLinear(torch.autograd.Function):
    """y = ((x @ W^T + bias + aux @ aux_w^T) [relu] + resid): forward product and input-gradient product on the tcgen05
    3xTF32 GEMM, weight / bias / aux_w gradients on the tcgen05 split-K weight-gradient kernel (all fp32-faithful).
    aux [M,3] (the vehicle scalars of FF_tour.0) carries no gradient."""

    @staticmethod
    def forward(ctx, x, W, bias, resid, relu, aux=None, aux_w=None):
        y = tc_gemm(x.detach(), W, bias, resid.detach() if resid is not None else None, relu, aux, aux_w)
        ctx.save_for_backward(x, W, y if relu else None, aux)
        ctx.flags = (bias is not None, resid is not None, relu)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W, y, aux = ctx.saved_tensors
        has_bias, has_resid, relu = ctx.flags
        gy = gy.contiguous()
        g = gy
        if relu:
            g = torch.empty_like(gy)
            _lib.check(_lib.lib().evrp_relu_mask(_lib.ptr(gy), _lib.ptr(y), _lib.ptr(g), gy.numel(), _lib.stream_ptr()),
                       "evrp_relu_mask")
        N, K = W.shape
        gx = gW = gb = gaw = None
        if ctx.needs_input_grad[0]:
            gx = tc_gemm(g, W.t()) if (K % 128 == 0 and N % 32 == 0) else g @ W
        want_b = has_bias and ctx.needs_input_grad[2]
        want_aw = aux is not None and ctx.needs_input_grad[6]
        if ctx.needs_input_grad[1] or want_b or want_aw:
            if N % 128 == 0 and K % 128 == 0:
                gW, gaux = tc_wgrad(g, x, want_b, aux if want_aw else None)
                if want_b:
                    gb = gaux[0]
                if want_aw:
                    gaw = gaux[1:].t()
            else:
                gW = g.t() @ x
                gb = g.sum(0) if want_b else None
                gaw = g.t() @ aux if want_aw else None
        return gx, gW, gb, (gy if has_resid else None), None, None, gaw


class FeedForwardTC(torch.autograd.Function):
    """y = relu(x W1^T + b1 [+ aux aux_w^T]) W2^T + b2 + resid  (nets/GraphEncoder.py:172-176 with its skip connection;
    FF_tour of nets/DRLModel.py:233-237 with the vehicle term) as ONE autograd node on the tcgen05 kernels: two forward
    GEMMs with fused epilogues; backward = input-gradient GEMM through W2 with the ReLU gate in its epilogue, the
    input-gradient GEMM through W1 and two split-K weight-gradient launches (bias and vehicle-weight rows included)."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, resid, aux=None, aux_w=None):
        f = tc_gemm(x.detach(), W1, b1, None, True, aux, aux_w)
        y = tc_gemm(f, W2, b2, resid.detach(), False)
        ctx.save_for_backward(x, W1, W2, f, aux)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W1, W2, f, aux = ctx.saved_tensors
        gy = gy.contiguous()
        gf = tc_gemm(gy, W2.t(), gate=f)                                   # d hidden, gated by the ReLU
        gW2, ga2 = tc_wgrad(gy, f, True)
        gx = tc_gemm(gf, W1.t()) if ctx.needs_input_grad[0] else None
        want_aw = aux is not None and ctx.needs_input_grad[7]
        gW1, ga1 = tc_wgrad(gf, x, True, aux if want_aw else None)
        return gx, gW1, ga1[0], gW2, ga2[0], gy, None, (ga1[1:].t() if want_aw else None)


class BatchNormTrain(torch.autograd.Function):
    """Train-mode BatchNorm1d(128) on the kernels of csrc/evrp_train.cu (nets/GraphEncoder.py:142-150): fp64 sufficient
    statistics [sum x, sum x^2, n]; when instances are sharded over ranks ONE all-reduce of these 257 doubles makes them the
    statistics of the GLOBAL batch (backward: one all-reduce of [sum g, sum g * xhat]).  Same outputs, gradients and
    running-statistics update as nn.BatchNorm1d on the concatenated batch."""

    @staticmethod
    def forward(ctx, x, weight, bias, bn, sync):
        import torch.distributed as dist
        L = _lib.lib()
        x = x.contiguous()
        M = x.shape[0]
        sums = torch.empty(257, device=x.device, dtype=torch.float64)
        _lib.check(L.evrp_bn_stats(_lib.ptr(x), M, _lib.ptr(sums), _lib.stream_ptr()), "evrp_bn_stats")
        if sync:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
        with torch.no_grad():
            bn.num_batches_tracked += 1
        save = torch.empty(257, device=x.device, dtype=torch.float32)
        y = torch.empty_like(x)
        w, b = weight.detach().contiguous(), bias.detach().contiguous()
        _lib.check(L.evrp_bn_apply(_lib.ptr(x), M, _lib.ptr(sums), _lib.ptr(w), _lib.ptr(b), float(bn.eps), float(m),
                                   _lib.ptr(bn.running_mean), _lib.ptr(bn.running_var), _lib.ptr(save), _lib.ptr(y),
                                   _lib.stream_ptr()), "evrp_bn_apply")
        ctx.save_for_backward(x, w, save)
        ctx.sync = sync
        return y

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist
        L = _lib.lib()
        x, w, save = ctx.saved_tensors
        g = g.contiguous()
        M = x.shape[0]
        sums = torch.empty(256, device=x.device, dtype=torch.float64)
        _lib.check(L.evrp_bn_backward_reduce(_lib.ptr(x), _lib.ptr(g), M, _lib.ptr(save), _lib.ptr(sums), _lib.stream_ptr()),
                   "evrp_bn_backward_reduce")
        local = sums.float()                                   # this rank's share: summed with the other gradients later
        if ctx.sync:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        gx = torch.empty_like(x)
        _lib.check(L.evrp_bn_backward_apply(_lib.ptr(x), _lib.ptr(g), M, _lib.ptr(save), _lib.ptr(w), _lib.ptr(sums), _lib.ptr(gx),
                                            _lib.stream_ptr()), "evrp_bn_backward_apply")
        return gx, local[128:], local[:128], None, None


class RouteContext(torch.autograd.Function):
    """(run_max [B,T,128], base [B,T,128]) of the teacher-forced decoder: running max of the visited embeddings (zeros before
    the first pick, nets/DRLModel.py:207-224) and fixed_ctx + emb[now] (:394) for all T recorded steps in one kernel; the
    backward scatters the gradients to the node holding the running max / the current node (csrc/evrp_train.cu)."""

    @staticmethod
    def forward(ctx, emb, fixed, pi, want_run_max):
        emb, fixed, pi = emb.contiguous(), fixed.contiguous(), pi.contiguous()
        B, S, d = emb.shape
        T = pi.shape[1]
        run_max = torch.empty(B, T, d, device=emb.device, dtype=torch.float32)
        base = torch.empty(B, T, d, device=emb.device, dtype=torch.float32)
        _lib.check(_lib.lib().evrp_route_context_forward(_lib.ptr(emb), _lib.ptr(fixed), _lib.ptr(pi), B, T, S, _lib.ptr(run_max),
                                                         _lib.ptr(base), _lib.stream_ptr()), "evrp_route_context_forward")
        ctx.save_for_backward(emb, pi)
        ctx.want_run_max = want_run_max
        return run_max, base

    @staticmethod
    def backward(ctx, g_rm, g_base):
        emb, pi = ctx.saved_tensors
        B, S, d = emb.shape
        T = pi.shape[1]
        g_emb = torch.empty_like(emb)
        g_fixed = torch.empty(B, d, device=emb.device, dtype=torch.float32)
        grm = g_rm.contiguous() if (ctx.want_run_max and g_rm is not None) else None
        if g_base is None:
            g_base = torch.zeros(B, T, d, device=emb.device)
        _lib.check(_lib.lib().evrp_route_context_backward(_lib.ptr(emb), _lib.ptr(pi), _lib.ptr(grm), _lib.ptr(g_base.contiguous()),
                                                          B, T, S, _lib.ptr(g_emb), _lib.ptr(g_fixed), _lib.stream_ptr()),
                   "evrp_route_context_backward")
        return g_emb, g_fixed, None, None


def self_attention_torch(qkv, Hh=8):
    """nets/GraphEncoder.py:81-102 in plain PyTorch: qkv [B,S,3*d] (Q | K | V, head-major) -> heads [B,S,d]
    (fp32 reference of csrc/evrp_attn_train.cu; materialises the [B,H,S,S] scores)"""
    B, S, d3 = qkv.shape
    dk = d3 // 3 // Hh
    x = qkv.view(B, S, 3, Hh, dk).permute(2, 0, 3, 1, 4)                               # [3,B,H,S,dk]
    attn = torch.softmax(torch.matmul(x[0], x[1].transpose(2, 3)) / math.sqrt(dk), -1)
    return torch.matmul(attn, x[2]).permute(0, 2, 1, 3).reshape(B, S, Hh * dk)


class SelfAttention(torch.autograd.Function):
    """encoder self-attention core of the training path on csrc/evrp_attn_train.cu (forward keeps the row
    log-sum-exp, backward recomputes the probabilities; no [B,H,S,S] tensor)"""

    @staticmethod
    def forward(ctx, qkv):
        qkv = qkv.contiguous()
        B, S, _ = qkv.shape
        heads = torch.empty(B, S, 128, device=qkv.device, dtype=torch.float32)
        lse = torch.empty(B, 8, S, device=qkv.device, dtype=torch.float32)
        _lib.check(_lib.lib().evrp_self_attention_forward(_lib.ptr(qkv), B, S, _lib.ptr(heads), _lib.ptr(lse),
                                                          _lib.stream_ptr()), "evrp_self_attention_forward")
        ctx.save_for_backward(qkv, heads, lse)
        return heads

    @staticmethod
    def backward(ctx, g):
        qkv, heads, lse = ctx.saved_tensors
        B, S, _ = qkv.shape
        gq = torch.empty_like(qkv)
        _lib.check(_lib.lib().evrp_self_attention_backward(_lib.ptr(qkv), _lib.ptr(heads), _lib.ptr(lse),
                                                           _lib.ptr(g.contiguous()), B, S, _lib.ptr(gq),
                                                           _lib.stream_ptr()), "evrp_self_attention_backward")
        return gq


class PointerLogProb(torch.autograd.Function):
    """ll[b,t] of the recorded actions from the step queries q and the encoder records kvl, forward and backward on the
    kernels of csrc/evrp_logp.cu (evrp_pointer_logp_forward / _backward).  Gradients flow to q (-> FF_tour, fixed
    context, emb[now]) and kvl (-> node projection, encoder); the trace tensors are constants."""

    @staticmethod
    def forward(ctx, q, kvl, mask_bits, pi, steps, tanh_clipping, temp):
        B, T, _ = q.shape
        S = kvl.shape[1]
        L = _lib.lib()
        ll = torch.empty(B, T, device=q.device, dtype=torch.float32)
        # the forward keeps each step's heads / softmax statistics for the backward (193 floats per step: ~120 MB at batch
        # 1024), which then skips its recompute passes over the K / V / logit-K records
        saved = torch.empty(B * T, L.evrp_pointer_logp_saved_floats(), device=q.device, dtype=torch.float32) \
            if (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) else None
        _lib.check(L.evrp_pointer_logp_forward(
            _lib.ptr(q), _lib.ptr(kvl), _lib.ptr(mask_bits), _lib.ptr(pi), _lib.ptr(steps), B, T, S, tanh_clipping, temp,
            _lib.ptr(ll), _lib.ptr(saved), _lib.stream_ptr()), "evrp_pointer_logp_forward")
        ctx.save_for_backward(q, kvl, mask_bits, pi, steps, saved)
        ctx.consts = (tanh_clipping, temp)
        return ll

    @staticmethod
    def backward(ctx, g):
        q, kvl, mask_bits, pi, steps, saved = ctx.saved_tensors
        B, T, _ = q.shape
        S = kvl.shape[1]
        gq = torch.empty_like(q)
        gk = torch.empty_like(kvl)
        _lib.check(_lib.lib().evrp_pointer_logp_backward(
            _lib.ptr(q), _lib.ptr(kvl), _lib.ptr(mask_bits), _lib.ptr(pi), _lib.ptr(steps), _lib.ptr(g.contiguous()),
            B, T, S, ctx.consts[0], ctx.consts[1], _lib.ptr(gq), _lib.ptr(gk), _lib.ptr(saved), _lib.stream_ptr()),
            "evrp_pointer_logp_backward")
        return gq, gk, None, None, None, None, None


def bits_to_mask(bits, S):
    """[...,4] int32 mask words -> [...,S] bool"""
    j = torch.arange(S, device=bits.device)
    words = bits.to(torch.int64) & 0xFFFFFFFF
    return ((words[..., j >> 5] >> (j & 31)) & 1).bool()


def teacher_forced_ll(model, emb, kvl, fixed, pi, mask_bits, veh, steps):
    """log-likelihood ll[B,T] of the recorded actions, differentiable w.r.t. emb / kvl / fixed and the decoder
    parameters.  Equivalent to running nets/DRLModel.py:259-276 step by step with the actions forced: given the
    actions, every step's inputs (running max, vehicle scalars, mask) are known up front, so the T steps are
    evaluated as one batch."""
    B, T = pi.shape
    S, d, Hh = emb.shape[1], emb.shape[2], model.n_heads
    policy = getattr(model, "policy", 0)
    if emb.is_cuda and d == 128 and policy == 0:
        # product path: running max + fixed_ctx + emb[now] of all T steps from ONE kernel (RouteContext), FF_tour on the tcgen05
        # GEMM with the vehicle term, bias and ReLU in its epilogue
        run_max, base = RouteContext.apply(emb, fixed, pi, True)
        W1 = model.FF_tour[0]
        q = FeedForwardTC.apply(run_max.view(B * T, d), W1.weight[:, 128:], W1.bias, model.FF_tour[2].weight,
                                model.FF_tour[2].bias, base.view(B * T, d), veh.reshape(B * T, 3),
                                folded_tour_weight(model)).view(B, T, d)
    else:
        now = torch.cat([torch.zeros_like(pi[:, :1]), pi[:, :-1]], 1)
        emb_now = emb.gather(1, now[..., None].expand(B, T, d))
        if policy == 1:                                                    # nets/AM.py:336-337,352-377
            q = fixed[:, None] + model.project_step_context(torch.cat([emb_now, veh[..., 0:1]], -1))
        else:
            idx = pi[..., None].expand(B, T, d)
            visited = emb.gather(1, idx)                                       # emb[pi_t]
            run_max = torch.cummax(visited, 1)[0]
            run_max = torch.cat([torch.zeros_like(run_max[:, :1]), run_max[:, :-1]], 1)    # zeros before first pick
            W1 = model.FF_tour[0]
            h1 = torch.relu(Fn.linear(veh, folded_tour_weight(model)) + Fn.linear(run_max, W1.weight[:, 128:]) + W1.bias)
            ctx = model.FF_tour[2](h1)
            q = (fixed[:, None] + ctx) + emb_now
    if q.is_cuda:                                                          # product path: hand-written forward + backward
        return PointerLogProb.apply(q.contiguous(), kvl.contiguous(), mask_bits.contiguous(), pi.contiguous(),
                                    steps.to(torch.int32).contiguous(), float(model.tanh_clipping), float(model.temp))
    # CPU tensors (tests only): the plain-PyTorch fp32 statement of the same op, the kernels' numerical reference
    return pointer_log_prob_torch(q, kvl, mask_bits, pi, steps, model.tanh_clipping, model.temp, Hh)


def pointer_log_prob_torch(q, kvl, mask_bits, pi, steps, tanh_clipping, temp, Hh=8):
    """nets/DRLModel.py:411-452 + :182-190 for all T steps at once in plain PyTorch (fp32 reference of
    csrc/evrp_logp.cu; materialises the [B,H,T,S] tensors the kernels avoid)."""
    B, T, d = q.shape
    S = kvl.shape[1]
    live = torch.arange(T, device=pi.device)[None] < steps[:, None]
    mask = bits_to_mask(mask_bits, S)
    mask = mask | (~live[..., None] & (torch.arange(S, device=pi.device) == 0))   # padded steps: depot only
    gK = kvl[..., :d].view(B, S, Hh, -1)
    gV = kvl[..., d:2 * d].view(B, S, Hh, -1)
    lK = kvl[..., 2 * d:]
    compat = torch.einsum("bthd,bshd->bhts", q.view(B, T, Hh, -1), gK) / math.sqrt(d // Hh)
    compat = compat.masked_fill(~mask[:, None], -math.inf)
    heads = torch.einsum("bhts,bshd->bthd", torch.softmax(compat, -1), gV).reshape(B, T, d)
    logits = torch.einsum("btd,bsd->bts", heads, lK) / math.sqrt(d)
    if tanh_clipping > 0:
        logits = torch.tanh(logits) * tanh_clipping
    logits = logits.masked_fill(~mask, -math.inf)
    log_p = torch.log_softmax(logits / temp, -1)
    ll = log_p.gather(2, pi[..., None])[..., 0]
    return torch.where(live, ll, torch.zeros_like(ll))
