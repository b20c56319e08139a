"""tests/ref_harness.py -- runs the UNMODIFIED reference (oracle/ref_model.py: models.model.PCNNet + metrics.loss.cd_loss_L1, byte-compiled
from /root/reference into oracle/_ref/refpy.zip) eagerly on the GPU as the full-size oracle of the -m gpu parity tests: its ATen operator chain
and its own Chamfer kernels (oracle/_ref/ref_chamfer3D.cubin).  Test infrastructure only."""
import torch


def available():
    from oracle import ref_chamfer as RC
    from oracle import ref_model as RM
    return RM.available() and RC.available()


def _pool_hook(store, key):
    def hook(mod, inp, out):
        # the selections the reference's VNMaxPool.forward just made (models/vn_layers.py:162-164), recomputed with the same ops
        x = inp[0].detach()
        d = mod.map_to_dir(x.transpose(1, -1)).transpose(1, -1)
        store[key] = (x * d).sum(2, keepdims=True).max(dim=-1, keepdim=False)[1].reshape(x.shape[0], -1)
    return hook


def reference_train_step(state_dict, p, c, R, tf32=False):
    """loads `state_dict` into the reference PCNNet (vn_pointnet + vn_foldingnet) on cuda, runs forward, the two L1-CD losses of
    train.py:151-160 and backward.  Returns dict(coarse, fine, loss, grads{name: tensor|None}, idx1, idx2, buffers{name: tensor})."""
    from oracle import ref_model as RM
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    try:
        net, ref = RM.build_pcnnet("cuda", backend="cuda")
        net.load_state_dict(state_dict)
        net.train()
        sel = {}
        h1 = net.encoder.maxpool1.register_forward_hook(_pool_hook(sel, "idx1"))
        h2 = net.encoder.maxpool2.register_forward_hook(_pool_hook(sel, "idx2"))
        coarse, fine = net(p, RM.Rotate(R))
        h1.remove()
        h2.remove()
        loss = ref.loss.cd_loss_L1(coarse, c) + ref.loss.cd_loss_L1(fine, c)
        loss.backward()
        torch.cuda.synchronize()
        out = dict(coarse=coarse.detach(), fine=fine.detach(), loss=float(loss.item()), idx1=sel["idx1"], idx2=sel["idx2"],
                   grads={n: (q.grad.detach().clone() if q.grad is not None else None) for n, q in net.named_parameters()},
                   buffers={n: b.detach().clone() for n, b in net.named_buffers()})
        del net, loss, coarse, fine
        torch.cuda.empty_cache()
        return out
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
