"""Model factory and checkpoint loader with the reference's signatures
(models/models.py:5-32, utils/utils.py:660-689).  Checkpoints are the reference's own files:
``torch.save({'net_param': dict, 'state_dict': OrderedDict})``."""
import torch

from . import _lib

NOT_YET = {}


def diffusion_models(task_model, net_param, **kwargs):
    """models/models.py:5-32."""
    if task_model == "NsDiff":
        from .nsdiff import NsDiff_model
        return NsDiff_model(net_param=net_param, train_model_select=kwargs["train_model_select"],
                            pretrain_f_path=net_param["pretrain_f_path"] if net_param.get("pretrain_f_path") else None,
                            pretrain_g_path=net_param["pretrain_g_path"] if net_param.get("pretrain_g_path") else None)
    if task_model == "NsDiff_spatial":
        from .nsdiff_spatial import NsDiff_model_spatial
        return NsDiff_model_spatial(net_param=net_param, train_model_select=kwargs["train_model_select"],
                                    pretrain_f_path=net_param["pretrain_f_path"] if net_param.get("pretrain_f_path") else None,
                                    pretrain_g_path=net_param["pretrain_g_path"] if net_param.get("pretrain_g_path") else None)
    if task_model == "NsDiff_model_variants":
        from .nsdiff import NsDiff_model_variants
        return NsDiff_model_variants(net_param=net_param, train_model_select=kwargs["train_model_select"])
    if task_model == "TMDM":
        from .tmdm import TMDM_model
        return TMDM_model(net_param=net_param)
    if task_model == "DiffusionTS":
        from .diffusionts import DiffusionTS_model
        return DiffusionTS_model(net_param=net_param)
    if task_model == "DiffSTG":
        from .diffstg import DiffSTG
        return DiffSTG(net_param=net_param)
    if task_model in NOT_YET:
        raise NotImplementedError(NOT_YET[task_model])
    raise ValueError("the definition  don't exit\n\tyou can define it before using it")


def _load_checkpoint(path):
    """{'net_param': dict, 'state_dict': OrderedDict} as utils/utils.py:611-622 saves it.  The reference full-unpickles
    (utils/utils.py:670); here the tensor-only loader is tried first -- the shipped checkpoints hold nothing but tensors,
    plain containers and a ``torch.device`` in net_param -- so that a third-party ``model_trained`` cannot run code on
    load.  A checkpoint that needs arbitrary pickled classes is refused unless UPD_ALLOW_PICKLE=1 opts in."""
    import os
    import pickle
    try:
        with torch.serialization.safe_globals([torch.device]):
            return torch.load(path, map_location="cpu", weights_only=True)
    except pickle.UnpicklingError as exc:
        if os.environ.get("UPD_ALLOW_PICKLE") != "1":
            raise RuntimeError("{} is not a tensor-only checkpoint ({}); set UPD_ALLOW_PICKLE=1 to full-unpickle it as "
                               "the reference does".format(path, str(exc).splitlines()[0])) from exc
        with open(path, "rb") as f:
            return torch.load(f, map_location=lambda storage, loc: storage, weights_only=False)


def drop_unused_temporal_tables(model, state_dict):
    """The reference builds TMDM's three embeddings as torch-timeseries ``DataEmbedding(c_in, d_model, 'fixed', 'h',
    dropout)`` (TMDM.py:90, tmdm_ns_transformer.py:53-56), which in that lineage registers a ``temporal_embedding`` of
    fixed sinusoidal lookup tables (``*.temporal_embedding.{hour,weekday,day,month}_embed.emb.weight``).  The hot path
    never passes time marks (x_mark is None: tmdm_adapter.py:123-124), so those tables are never read; the modules here
    do not hold them.  They are dropped from a checkpoint -- and only they -- so that the strict load still checks every
    tensor the path uses.  (torch-timeseries is not vendored and no TMDM checkpoint ships: the key names are the
    lineage's, unverified.)"""
    own = set(model.state_dict().keys())
    return {k: v for k, v in state_dict.items() if k in own or ".temporal_embedding." not in k}


def load_diffusion_model(path, device, infer_para=None, dataparallel=True, **kwargs):
    """utils/utils.py:660-689: full-pickle load, ``infer_para`` merged before construction (so it can
    change n_z_samples / parallel_sample / diffusion_steps), ``module.`` prefixes stripped,
    ``net_param['device']`` overwritten, strict state-dict load.  -> (model, loaded_net_param)."""
    device = _lib.require_cuda(device)
    state = _load_checkpoint(path)
    loaded_net_param = state["net_param"]
    if infer_para is not None:
        loaded_net_param.update(infer_para)
    loaded_state_dict = state["state_dict"]
    # the reference keeps the ``module.`` prefix when several GPUs are visible because it would wrap the model in
    # DataParallel; this build runs one process per GPU and never wraps, so the prefix always goes
    loaded_state_dict = {k.replace("module.", ""): v for k, v in loaded_state_dict.items()}
    loaded_net_param["device"] = device
    model = diffusion_models(task_model=loaded_net_param["task_model"], net_param=loaded_net_param,
                             train_model_select=kwargs["train_model_select"]).to(device)
    model.load_state_dict(drop_unused_temporal_tables(model, loaded_state_dict), strict=True)
    model = model.to(device)
    return model, loaded_net_param
