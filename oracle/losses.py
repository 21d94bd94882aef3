"""CPU oracle of the training losses.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates model/loss.py (L2Loss 15-43, GaussianDkl 46-66, SynthParamsLoss 73-183) and the step-body sum of
train.py:222-246.  `synth_params_loss` is the vectorised equivalent of the reference's per-row / per-group Python
loops; oracle/make_golden.py asserts it equals the reference class on the seeded inputs.  Unlike the reference
(loss.py:134-135) it does not mutate its inputs.
"""
import torch
from torch.nn import functional as F


def l2_loss(inferred, target, contents_average=False, batch_average=True):       # loss.py:15-43
    loss = torch.sum(torch.square(inferred - target))
    if batch_average:
        loss = loss / inferred.shape[0]
    if contents_average:
        loss = loss / inferred[0, :].numel()
    return loss


def useless_masks(idx_helper, u_in):
    """[B, n_num] and [B, n_groups] boolean masks of outputs whose operator has target output level < 1e-3
    (data/preset.py:247-283, loss.py:120-126)."""
    t = idx_helper.device_tables() if hasattr(idx_helper, 'device_tables') else _tables_from_reference(idx_helper)
    def mask(vol_cols):
        vol = torch.as_tensor(vol_cols, dtype=torch.long, device=u_in.device)
        m = torch.zeros(u_in.shape[0], len(vol_cols), dtype=torch.bool, device=u_in.device)
        has = vol >= 0
        m[:, has] = u_in[:, vol[has]] < 1e-3
        return m
    return t, mask(t['num_vol_col']), mask(t['grp_vol_col'])


def _tables_from_reference(ref_helper):
    """Build the flat tables from a REFERENCE PresetIndexesHelper by probing get_useless_learned_params_indexes."""
    import numpy as np
    num_cols = ref_helper.get_numerical_learnable_indexes()
    groups = ref_helper.get_categorical_learnable_indexes()
    L = ref_helper.learnable_preset_size
    num_vol = {c: -1 for c in num_cols}
    grp_vol = {g[0]: -1 for g in groups}
    for op in range(6):
        vol = ref_helper.full_to_learnable[31 + 22 * op]
        if not isinstance(vol, int):
            continue
        probe = torch.ones(L)
        probe[vol] = 0.0
        n, c = ref_helper.get_useless_learned_params_indexes(probe)
        for col in n:
            num_vol[col] = vol
        for col in c:
            grp_vol[col] = vol
    i32 = lambda a: np.asarray(a, dtype=np.int32)
    return dict(num_cols=i32(num_cols), num_vol_col=i32([num_vol[c] for c in num_cols]),
                grp_start=i32([g[0] for g in groups]), grp_len=i32([len(g) for g in groups]),
                grp_vol_col=i32([grp_vol[g[0]] for g in groups]))


def synth_params_loss(idx_helper, u_out, u_in, normalize_losses=True, categorical_loss_factor=0.2,
                      prevent_useless_params_loss=True, cat_bce=False, cat_softmax=True, cat_softmax_t=0.2):
    """loss.py:117-183.  Defaults are the ones train.py:110-116 derives from config.py (CCE, softmax T=0.2)."""
    t, useless_num, useless_grp = useless_masks(idx_helper, u_in)
    if not prevent_useless_params_loss:
        useless_num[:] = False
        useless_grp[:] = False
    B = u_in.shape[0]
    num_cols = torch.as_tensor(t['num_cols'], dtype=torch.long, device=u_in.device)
    num_loss = 0.0
    if len(num_cols) > 0:
        keep = (~useless_num).to(u_out.dtype)
        o, i = u_out[:, num_cols] * keep, u_in[:, num_cols] * keep           # loss.py:128-135 zeroes both sides
        num_loss = F.mse_loss(o, i, reduction='mean') if normalize_losses else l2_loss(o, i)
    cat_loss = 0.0
    n_groups = len(t['grp_start'])
    for g in range(n_groups):
        s, n = int(t['grp_start'][g]), int(t['grp_len'][g])
        useful = ~useless_grp[:, g]
        q = u_out[:, s:s + n][useful]
        target = u_in[:, s:s + n][useful]
        if not cat_bce:
            if cat_softmax:
                q = torch.softmax(q / cat_softmax_t, dim=1)                  # loss.py:167-168
            picked = q[target.bool()]                                         # loss.py:170
            cat_loss = cat_loss - torch.sum(torch.log(picked)) / int(useful.sum())   # loss.py:172
        else:
            cat_loss = cat_loss + F.binary_cross_entropy(q, target, reduction='mean') / 8.0
    if n_groups > 0 and normalize_losses:
        cat_loss = cat_loss / n_groups                                        # loss.py:180-181
    return num_loss + cat_loss * categorical_loss_factor                      # loss.py:183


def train_step_losses(ext_model, x_in, v_in, sample_info, noise, beta, normalize_losses=True,
                      cat_softmax_t=0.2, params_reg_softmax=False):
    """Forward + the four loss terms of train.py:209-241 (flow_input_loss is zero for the default 'bn'
    regularisation, train.py:235-239).  Returns (outputs dict, losses dict, total)."""
    z0_mu_logvar, z0, zK, logdet, x_out = ext_model(x_in, sample_info, noise)
    v_out = ext_model.reg_model(zK, noise)
    recons = F.mse_loss(x_out, x_in, reduction='mean') if normalize_losses else l2_loss(x_out, x_in)
    lat = ext_model.latent_loss(z0_mu_logvar, z0, zK, logdet)
    cont = synth_params_loss(ext_model.idx_helper, v_out, v_in, normalize_losses=normalize_losses,
                             cat_softmax=not params_reg_softmax, cat_softmax_t=cat_softmax_t)
    total = recons + beta * lat + cont
    outs = dict(z0_mu_logvar=z0_mu_logvar, z0=z0, zK=zK, logdet=logdet, x_out=x_out, v_out=v_out)
    return outs, dict(recons=recons, latent=lat, controls=cont), total
