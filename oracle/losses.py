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


# ------------------------------------------------------------------------------------------------ monitoring metrics (train.py:232-233)
def quantized_numerical_params_loss(idx_helper, u_out, u_in, l1=False):
    """model/loss.py:217-261 (QuantizedNumericalParamsLoss.__call__ with numerical_loss = nn.MSELoss() / nn.L1Loss(), all parameters)."""
    cols_in, cols_out = [], []
    for vst_idx, learn_idx in idx_helper.num_idx_learned_as_num.items():          # loss.py:231-243
        o = u_out[:, learn_idx].detach().clone()
        card = idx_helper.vst_param_cardinals[vst_idx]
        if card > 0:
            o = torch.round(o * (card - 1.0)) / (card - 1.0)
        cols_in.append(u_in[:, learn_idx].detach())
        cols_out.append(o)
    for vst_idx, learn_indexes in idx_helper.num_idx_learned_as_cat.items():      # loss.py:245-255
        card = len(learn_indexes)
        cols_in.append(torch.argmax(u_in[:, learn_indexes], dim=-1).float() / (card - 1.0))
        cols_out.append(torch.argmax(u_out[:, learn_indexes], dim=-1).float() / (card - 1.0))
    a, b = torch.stack(cols_out, 1), torch.stack(cols_in, 1)
    return F.l1_loss(a, b) if l1 else F.mse_loss(a, b)


def categorical_params_accuracy(idx_helper, u_out, u_in, reduce=True, percentage_output=True):
    """model/loss.py:283-315."""
    acc = {}
    for vst_idx, learn_idx in idx_helper.cat_idx_learned_as_num.items():          # loss.py:289-299
        card = idx_helper.vst_param_cardinals[vst_idx]
        t = torch.round(u_in[:, learn_idx] * (card - 1.0)).to(torch.int32)
        o = torch.round(u_out[:, learn_idx] * (card - 1.0)).to(torch.int32)
        acc[vst_idx] = (t == o).count_nonzero().item() / t.numel()
    for vst_idx, learn_indexes in idx_helper.cat_idx_learned_as_cat.items():      # loss.py:301-307
        t, o = torch.argmax(u_in[:, learn_indexes], dim=-1), torch.argmax(u_out[:, learn_indexes], dim=-1)
        acc[vst_idx] = (t == o).count_nonzero().item() / t.numel()
    if percentage_output:
        acc = {k: v * 100.0 for k, v in acc.items()}
    return float(sum(acc.values()) / len(acc)) if reduce else acc


def flow_params_loss(latent_flow_inverse, reg_flow_inverse, z0_mu_logvar, v_target):
    """model/loss.py:333-346 with utils/probability.py:21-29."""
    import numpy as np
    z_K, ld_u = reg_flow_inverse(v_target)
    z_0, ld_t = latent_flow_inverse(z_K)
    mu, lv = z0_mu_logvar[:, 0, :], z0_mu_logvar[:, 1, :]
    log_q = -0.5 * (z_0.shape[1] * np.log(2 * np.pi) + torch.sum(lv + (z_0 - mu) ** 2 / torch.exp(lv), dim=1))
    return -torch.mean(log_q + ld_t + ld_u) / 1000.0


def learnable_to_full(idx_helper, learnable, default_values):
    """data/preset.py:350-369 (PresetsParams.get_full from learnable presets)."""
    full = -0.1 * torch.ones((learnable.shape[0], idx_helper.full_preset_size))
    for vst_idx, learn in enumerate(idx_helper.full_to_learnable):
        if idx_helper.vst_param_learnable_model[vst_idx] is None:
            if vst_idx in default_values:
                full[:, vst_idx] = default_values[vst_idx]
        elif isinstance(learn, int):
            full[:, vst_idx] = learnable[:, learn]
        else:
            full[:, vst_idx] = torch.argmax(learnable[:, learn], dim=-1) / (idx_helper.vst_param_cardinals[vst_idx] - 1.0)
    return full


def spectrogram_stats(specs):
    """data/abstractbasedataset.py:357-386: per-item (min, max, mean, torch.var) and the data-set summary of :357-360."""
    import numpy as np
    per = np.asarray([[s.min().item(), s.max().item(), torch.mean(s, dim=(0, 1)).item(), torch.var(s).item()] for s in specs], dtype=np.float64)
    return per, dict(min=per[:, 0].min(), max=per[:, 1].max(), mean=per[:, 2].mean(), std=float(np.sqrt(per[:, 3].mean())))
