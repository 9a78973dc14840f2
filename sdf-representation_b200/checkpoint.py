"""Checkpoint layout of the reference, byte-for-byte compatible pickles (``executor/executor.py:237-345``):
``models/best_model.pkl`` = {epoch, loss_per_epoch, best_val_loss, val_loss_per_epoch, model_state_dict,
optimizer_state_dict}; ``models/model_epoch{E}.pkl`` = {epoch, model_state_dict}; plain ``pickle`` (not torch.save);
``module.`` prefixes added / stripped as ``model_device_handler`` does; CPU hosts load CUDA pickles through the
``CPU_Unpickler`` of ``utils/pickling.py:7-11``."""
import glob
import io
import os
import pickle
from collections import OrderedDict

import torch


class CPU_Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "torch.storage" and name == "_load_from_bytes":
            return lambda b: torch.load(io.BytesIO(b), map_location="cpu")
        return super().find_class(module, name)


def save_model(model, optimizer, loss_per_epoch, epoch, best_val_loss, val_loss_per_epoch, save_path, best=False):
    if best:
        data = {
            "epoch": epoch,
            "loss_per_epoch": loss_per_epoch,
            "best_val_loss": best_val_loss,
            "val_loss_per_epoch": val_loss_per_epoch,
            "model_state_dict": model.state_dict(),
            "optimizer_state_dict": optimizer.state_dict(),
        }
        path = os.path.join(save_path, "best_model.pkl")
    else:
        data = {"epoch": epoch, "model_state_dict": model.state_dict()}
        path = os.path.join(save_path, f"model_epoch{epoch}.pkl")
    with open(path, "wb") as f:
        pickle.dump(data, f)
    return path


def _load_pickle(path):
    with open(path, "rb") as f:
        if not torch.cuda.is_available():
            return CPU_Unpickler(f).load()
        return pickle.load(f)


def model_device_handler(model, model_state_dict):
    wrapped = isinstance(model, torch.nn.DataParallel)
    new_state = OrderedDict()
    for k, v in model_state_dict.items():
        if wrapped:
            name = k if k.startswith("module.") else "module." + k
        else:
            name = k[7:] if k.startswith("module.") else k
        new_state[name] = v
    model.load_state_dict(new_state)
    return model


def load_model(model, optimizer, save_path, best=False):
    if best:
        data = _load_pickle(os.path.join(save_path, "best_model.pkl"))
        model = model_device_handler(model, data["model_state_dict"])
        optimizer.load_state_dict(data["optimizer_state_dict"])
        return model, optimizer, data["epoch"] + 1, data["loss_per_epoch"], data["best_val_loss"], data["val_loss_per_epoch"]
    files = glob.glob(save_path + "/model_epoch*.pkl")
    if not files:
        raise FileNotFoundError(f"no model_epoch*.pkl under {save_path}")
    files.sort(key=os.path.getmtime, reverse=True)  # newest by mtime, executor/executor.py:283-287
    data = _load_pickle(files[0])
    model = model_device_handler(model, data["model_state_dict"])
    return model, data["epoch"] + 1


def to_compatible_keys(state_dict):
    """lin{i}.* -> layers.{i}.* (ImplicitNetCompatible, utils/inference_conversion.py:39-45) for TorchScript/ONNX users."""
    out = OrderedDict()
    for k, v in state_dict.items():
        k = k[7:] if k.startswith("module.") else k
        if k.startswith("lin"):
            idx, rest = k[3:].split(".", 1)
            out[f"layers.{idx}.{rest}"] = v
        else:
            out[k] = v
    return out


def export_flat(model, path):
    """Write the network as an ``SDFW0001`` flat binary for the native query tool (``bin/sdf_query``, the equivalent of
    the reference's ``utils/inference_conversion.py`` -> ``ops/DeepTrace`` / ``ops/conversion_test`` hand-off):
    magic, int32 d_in / hidden / num_hidden / skip_layer, float32 beta, then per layer weight [out,in] and bias."""
    import struct

    net = model.module if hasattr(model, "module") else model
    n_lin = net.num_layers - 1
    hidden = getattr(net, "lin0").weight.shape[0]
    skip = int(net.skip_in[0]) if len(net.skip_in) else 0
    beta = float(getattr(net, "beta", 0.0) or 0.0)
    with open(path, "wb") as f:
        f.write(b"SDFW0001")
        f.write(struct.pack("<4if", int(getattr(net, "lin0").weight.shape[1]), int(hidden), int(n_lin - 1), skip, beta))
        for i in range(n_lin):
            lin = getattr(net, f"lin{i}")
            f.write(lin.weight.detach().float().cpu().contiguous().numpy().tobytes())
            f.write(lin.bias.detach().float().cpu().contiguous().numpy().tobytes())
    return path
