"""Preset index tables consumed by the regression head and the synth-parameter loss (host logic, no kernels).

Mirrors the interface of the reference's `data.preset.PresetIndexesHelper` (data/preset.py:23-283) for the
members the hot path reads: `full_to_learnable`, `learnable_to_full`, `learnable_preset_size`,
`vst_param_learnable_model`, `vst_param_cardinals`, `get_numerical_learnable_indexes()`,
`get_categorical_learnable_indexes()`, `get_useless_learned_params_indexes(preset_GT)`.  The product modules
duck-type on these, so the reference's own helper (built from its DexedDataset) can be passed instead.

`DexedLearnableLayout` replaces the part of data/dexeddataset.py:79-167 that decides, from the static Dexed
parameter tables (synth/dexed.py:385-452), which VST parameters are learnable and how ('num' / 'cat').  It needs
neither the preset SQLite DB nor the RenderMan plugin, which are absent from the reference tree.

`device_tables()` packs the same information into flat int32 arrays for the CUDA loss kernels
(pgv_synth_params_loss_fwd_bwd, include/pgv.h).
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

DEXED_PARAMS_COUNT = 155          # synth/dexed.py (presets_mat has 155 columns)
_OP_FIRST, _OP_STRIDE, _N_OPS = 23, 22, 6
# offsets inside one operator block of 22 VST params (block starts at 23 + 22*op)
_OP_OUTPUT_LEVEL, _OP_SWITCH = 8, 21
_OP_CARD = {9: 2, 10: 32, 12: 15, 16: 4, 17: 4, 18: 8, 19: 4, 20: 8, 21: 2}   # synth/dexed.py:398-418
_OP_CATEGORICAL = (9, 16, 17, 21)                                               # synth/dexed.py:447-451
_GENERAL_CARD = {4: 32, 5: 8, 6: 2, 11: 2, 12: 6, 14: 8}                        # synth/dexed.py:388-397
_GENERAL_CATEGORICAL = (4, 6, 11, 12)                                           # synth/dexed.py:446


def dexed_param_cardinality(vst_idx: int) -> int:
    """Number of discrete values of a Dexed VST parameter, -1 if treated as continuous (synth/dexed.py:385-422)."""
    if vst_idx < _OP_FIRST:
        return _GENERAL_CARD.get(vst_idx, -1)
    return _OP_CARD.get((vst_idx - _OP_FIRST) % _OP_STRIDE, -1)


def dexed_categorical_vst_params() -> List[int]:
    out = list(_GENERAL_CATEGORICAL)
    for op in range(_N_OPS):
        out += [_OP_FIRST + _OP_STRIDE * op + o for o in _OP_CATEGORICAL]
    return out


def dexed_numerical_vst_params() -> List[int]:
    # synth/dexed.py:425-442 lists them explicitly; they are exactly the non-categorical ones
    cat = set(dexed_categorical_vst_params())
    return [i for i in range(DEXED_PARAMS_COUNT) if i not in cat]


class DexedLearnableLayout:
    """Dataset stand-in exposing exactly what `PresetIndexesHelper(dataset)` reads (data/preset.py:55-86)."""
    synth_name = "Dexed"

    def __init__(self, vst_params_learned_as_categorical: Optional[str] = 'all<=32',
                 operators: Sequence[int] = (1, 2, 3, 4, 5, 6), constant_filter_and_tune_params: bool = True,
                 algos: Sequence[int] = ()):
        self.algos = list(algos)
        self.total_nb_params = DEXED_PARAMS_COUNT
        self.preset_param_names = ['p{}'.format(i) for i in range(DEXED_PARAMS_COUNT)]
        learnable = set(range(DEXED_PARAMS_COUNT))
        self.params_default_values = {}
        card = np.asarray([dexed_param_cardinality(i) for i in range(DEXED_PARAMS_COUNT)])
        if constant_filter_and_tune_params:              # data/dexeddataset.py:80-82, 131-137
            for i, v in zip((0, 1, 2, 3, 13), (1.0, 0.0, 1.0, 0.5, 0.5)):
                learnable.discard(i)
                card[i] = 1
                self.params_default_values[i] = v
        for op in range(_N_OPS):                         # data/dexeddataset.py:83-89, 126-130
            first = _OP_FIRST + _OP_STRIDE * op
            if (op + 1) not in operators:
                learnable -= set(range(first, first + 21))
            learnable.discard(first + _OP_SWITCH)
            card[first + _OP_SWITCH] = 1
            self.params_default_values[first + _OP_SWITCH] = 1.0 if (op + 1) in operators else 0.0
        if len(self.algos) == 1:
            learnable.discard(4)
            self.params_default_values[4] = (self.algos[0] - 1) / 31.0
        if len(self.algos) > 0:
            card[4] = len(self.algos)
        self._card = card
        self.learnable_params_idx = sorted(learnable)
        self.numerical_vst_params = dexed_numerical_vst_params()
        self.categorical_vst_params = dexed_categorical_vst_params()
        threshold = None
        if vst_params_learned_as_categorical is not None:
            if vst_params_learned_as_categorical.startswith('all<='):
                threshold = int(vst_params_learned_as_categorical.replace('all<=', ''))
            elif vst_params_learned_as_categorical != 'vst_cat':
                raise ValueError(vst_params_learned_as_categorical)
        model = []
        for i in range(DEXED_PARAMS_COUNT):              # data/dexeddataset.py:148-167
            if i not in learnable:
                model.append(None)
            elif vst_params_learned_as_categorical is None:
                model.append('num')
            elif i in self.categorical_vst_params:
                model.append('cat')
            elif threshold is not None and 1 < card[i] <= threshold:
                model.append('cat')
            else:
                model.append('num')
        self.vst_param_learnable_model = model
        self.preset_indexes_helper = PresetIndexesHelper(self)

    def get_preset_param_cardinality(self, idx, learnable_representation=True):
        return int(self._card[idx])


class PresetIndexesHelper:
    """VST-index <-> learnable-column translation tables (interface of data/preset.py:23-283)."""

    def __init__(self, dataset=None, nb_params=None):
        if dataset is None:
            assert nb_params is not None
            self.synth_name = "generic_synth"
            self._vst_param_learnable_model = ['num'] * nb_params
            self._param_cardinals = [-1] * nb_params
            self._numerical_vst_params = list(range(nb_params))
            self._categorical_vst_params = []
            self._param_names = ['param'] * nb_params
        else:
            assert nb_params is None
            self.synth_name = dataset.synth_name
            self._vst_param_learnable_model = list(dataset.vst_param_learnable_model)
            self._param_cardinals = [dataset.get_preset_param_cardinality(i, learnable_representation=True)
                                     for i in range(dataset.total_nb_params)]
            self._numerical_vst_params = list(dataset.numerical_vst_params)
            self._categorical_vst_params = list(dataset.categorical_vst_params)
            self._param_names = list(dataset.preset_param_names)
        self._is_dexed = self.synth_name.lower() == "dexed"
        self._full_to_learnable, self._learnable_to_full = [], []
        col = 0
        for vst_idx, m in enumerate(self._vst_param_learnable_model):
            if m is None:
                self._full_to_learnable.append(None)
            elif m == 'num':
                self._full_to_learnable.append(col)
                self._learnable_to_full.append(vst_idx)
                col += 1
            elif m == 'cat':
                n = self._param_cardinals[vst_idx]
                self._full_to_learnable.append(list(range(col, col + n)))
                self._learnable_to_full += [vst_idx] * n
                col += n
            else:
                raise ValueError("Unknown param learning model '{}'".format(m))
        self._learnable_preset_size = col

        def _select(vst_params, model):
            return {i: self._full_to_learnable[i] for i in vst_params if self._vst_param_learnable_model[i] == model}
        self.cat_idx_learned_as_num = _select(self._categorical_vst_params, 'num')
        self.cat_idx_learned_as_cat = _select(self._categorical_vst_params, 'cat')
        self.num_idx_learned_as_num = _select(self._numerical_vst_params, 'num')
        self.num_idx_learned_as_cat = _select(self._numerical_vst_params, 'cat')

    # - - - same read-only surface as the reference helper - - -
    full_preset_size = property(lambda self: len(self._full_to_learnable))
    vst_param_names = property(lambda self: self._param_names)
    numerical_vst_params = property(lambda self: self._numerical_vst_params)
    categorical_vst_params = property(lambda self: self._categorical_vst_params)
    vst_param_learnable_model = property(lambda self: self._vst_param_learnable_model)
    vst_param_cardinals = property(lambda self: self._param_cardinals)
    full_to_learnable = property(lambda self: self._full_to_learnable)
    learnable_to_full = property(lambda self: self._learnable_to_full)
    learnable_preset_size = property(lambda self: self._learnable_preset_size)

    def get_numerical_learnable_indexes(self) -> List[int]:
        return [self._full_to_learnable[i] for i, m in enumerate(self._vst_param_learnable_model) if m == 'num']

    def get_categorical_learnable_indexes(self) -> List[List[int]]:
        return [self._full_to_learnable[i] for i, m in enumerate(self._vst_param_learnable_model) if m == 'cat']

    def get_useless_learned_params_indexes(self, preset_GT: torch.Tensor):
        """Per-row host version of data/preset.py:247-283 (kept for API parity; the training path uses the
        device tables below instead of calling this once per row)."""
        if not self._is_dexed:
            return [], []
        num_cols, cat_first_cols = [], []
        for op in range(_N_OPS):
            vol = self._full_to_learnable[_OP_FIRST + _OP_STRIDE * op + _OP_OUTPUT_LEVEL]
            if vol is None:
                continue
            if not isinstance(vol, int):
                raise NotImplementedError("Dexed Operator output volume learned as categorical")
            if preset_GT[vol].item() < 1e-3:
                for off in range(21):
                    if off == _OP_OUTPUT_LEVEL:
                        continue
                    learn = self._full_to_learnable[_OP_FIRST + _OP_STRIDE * op + off]
                    if isinstance(learn, int):
                        num_cols.append(learn)
                    elif isinstance(learn, list):
                        cat_first_cols.append(learn[0])
        return num_cols, cat_first_cols

    def device_tables(self):
        """Flat tables for the CUDA loss kernels.

        Returns a dict of int32 numpy arrays:
          num_cols[n_num]            learnable column of each numerical output
          num_vol_col[n_num]         column whose target < 1e-3 disables this output, or -1
          grp_start[n_grp], grp_len[n_grp]   contiguous column range of each categorical group
          grp_vol_col[n_grp]         column whose target < 1e-3 removes the row from this group's CE, or -1
        """
        def volume_col(vst_idx):
            if not self._is_dexed or vst_idx < _OP_FIRST:
                return -1
            op, off = divmod(vst_idx - _OP_FIRST, _OP_STRIDE)
            if off in (_OP_OUTPUT_LEVEL, _OP_SWITCH):
                return -1
            vol = self._full_to_learnable[_OP_FIRST + _OP_STRIDE * op + _OP_OUTPUT_LEVEL]
            if vol is None:
                return -1
            if not isinstance(vol, int):
                raise NotImplementedError("Dexed Operator output volume learned as categorical")
            return vol
        num_cols, num_vol, g_start, g_len, g_vol = [], [], [], [], []
        for vst_idx, m in enumerate(self._vst_param_learnable_model):
            learn = self._full_to_learnable[vst_idx]
            if m == 'num':
                num_cols.append(learn)
                num_vol.append(volume_col(vst_idx))
            elif m == 'cat':
                g_start.append(learn[0])
                g_len.append(len(learn))
                g_vol.append(volume_col(vst_idx))
        i32 = lambda a: np.asarray(a, dtype=np.int32)
        return dict(num_cols=i32(num_cols), num_vol_col=i32(num_vol), grp_start=i32(g_start), grp_len=i32(g_len),
                    grp_vol_col=i32(g_vol))


class tables_from_foreign_helper:
    """Adapter giving `device_tables()` to a helper that does not have it (e.g. the reference's own
    data.preset.PresetIndexesHelper): the operator -> affected-columns map is recovered by probing
    `get_useless_learned_params_indexes` with one silent operator at a time."""

    def __init__(self, helper):
        self.helper = helper

    def device_tables(self):
        h = self.helper
        num_cols = h.get_numerical_learnable_indexes()
        groups = h.get_categorical_learnable_indexes()
        num_vol = {c: -1 for c in num_cols}
        grp_vol = {g[0]: -1 for g in groups}
        for op in range(_N_OPS):
            vst = _OP_FIRST + _OP_STRIDE * op + _OP_OUTPUT_LEVEL
            vol = h.full_to_learnable[vst] if vst < len(h.full_to_learnable) else None
            if not isinstance(vol, int):
                continue
            probe = torch.ones(h.learnable_preset_size)
            probe[vol] = 0.0
            n, c = h.get_useless_learned_params_indexes(probe)
            for col in n:
                num_vol[col] = vol
            for col in c:
                grp_vol[col] = vol
        i32 = lambda a: np.asarray(a, dtype=np.int32)
        return dict(num_cols=i32(num_cols), num_vol_col=i32([num_vol[c] for c in num_cols]),
                    grp_start=i32([g[0] for g in groups]), grp_len=i32([len(g) for g in groups]),
                    grp_vol_col=i32([grp_vol[g[0]] for g in groups]))


PM_NUM_AS_NUM, PM_NUM_AS_CAT, PM_CAT_AS_NUM, PM_CAT_AS_CAT, PM_NONE = range(5)


def metric_tables(idx_helper, limited_vst_params_indexes=None, default_values=None):
    """Per-VST-parameter tables of the device metric / conversion kernels (pgv.h, pgv_preset_metrics / pgv_learnable_to_full), built
    from the read-only surface shared by this package's PresetIndexesHelper and the reference's (data/preset.py:23-283):
    kind (0 numerical->numerical, 1 numerical->one-hot, 2 categorical->numerical, 3 categorical->one-hot, 4 not learnable or
    excluded by `limited_vst_params_indexes`), col / len (learnable column range), card, fill (default value or -0.1)."""
    n = idx_helper.full_preset_size
    kind, col, ln, card, fill = (np.full(n, PM_NONE, np.int32), np.zeros(n, np.int32), np.ones(n, np.int32), np.zeros(n, np.int32),
                                 np.full(n, -0.1, np.float32))
    numerical = set(idx_helper.numerical_vst_params)
    for vst_idx, learn in enumerate(idx_helper.full_to_learnable):
        card[vst_idx] = int(idx_helper.vst_param_cardinals[vst_idx])
        if default_values is not None and vst_idx in default_values:
            fill[vst_idx] = default_values[vst_idx]
        if learn is None or (limited_vst_params_indexes is not None and vst_idx not in limited_vst_params_indexes):
            continue
        as_cat = not isinstance(learn, int)
        col[vst_idx] = learn[0] if as_cat else learn
        ln[vst_idx] = len(learn) if as_cat else 1
        if vst_idx in numerical:
            kind[vst_idx] = PM_NUM_AS_CAT if as_cat else PM_NUM_AS_NUM
        else:
            kind[vst_idx] = PM_CAT_AS_CAT if as_cat else PM_CAT_AS_NUM
    return dict(kind=kind, col=col, len=ln, card=card, fill=fill)


def learnable_to_full_presets(idx_helper, learnable_presets: torch.Tensor, default_values: dict) -> torch.Tensor:
    """Inference tail (data/preset.py:350-369).  CUDA tensors go through pgv_learnable_to_full (one launch; per-group argmax on the
    device); CPU tensors through the vectorised torch code below (host-side tooling, e.g. writing presets to disk)."""
    if learnable_presets.is_cuda:
        from ..model import ops
        return ops.learnable_to_full(learnable_presets, ops.MetricTables(idx_helper, default_values=default_values))
    return _learnable_to_full_presets_host(idx_helper, learnable_presets, default_values)


def _learnable_to_full_presets_host(idx_helper, learnable_presets: torch.Tensor, default_values: dict) -> torch.Tensor:
    """Inference tail (data/preset.py:350-369): per-group argmax / (n-1) for categorical groups, copy for
    numerical columns, defaults (else -0.1) for non-learnable VST parameters.  Vectorised, runs on the tensor's device."""
    B = learnable_presets.shape[0]
    full = torch.full((B, idx_helper.full_preset_size), -0.1, device=learnable_presets.device)
    for vst_idx, learn in enumerate(idx_helper.full_to_learnable):
        if learn is None:
            if vst_idx in default_values:
                full[:, vst_idx] = default_values[vst_idx]
        elif isinstance(learn, int):
            full[:, vst_idx] = learnable_presets[:, learn]
        else:
            n = idx_helper.vst_param_cardinals[vst_idx]
            full[:, vst_idx] = torch.argmax(learnable_presets[:, learn[0]:learn[-1] + 1], dim=-1) / (n - 1.0)
    return full
