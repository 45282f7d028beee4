"""On-disk format either side of the path: loading a reference ``nerf.pth`` into this repo's
``NeRFNetwork`` and writing one back.

What the reference stores.  ``train.py:307`` pickles the whole unwrapped module (``torch.save(nerf, 'nerf.pth')``) and
``train.py:472`` reads it back with ``torch.load``; a ``state_dict`` of that module has these entries:

    sigma_net.params     fp32 [n_mlp_sigma + n_table]   tcnn.NetworkWithInputEncoding (nerf/network.py:23-26): the MLP
                                                        matrices in layer order, each row-major [out, in] (out padded
                                                        to 16 on the last), followed by the hash table, level-major,
                                                        n_features_per_level floats per entry
    encoder_dir.params   fp32 [0]                       tcnn.Encoding, SphericalHarmonics (:29-32): no parameters
    color_net.params     fp32 [n_mlp_color]             tcnn.Network (:34-37): input padded 31 -> 32.  tiny-cuda-nn feeds 1.0
                                                        into the pad (Identity-encoding padding), so column 31 of the first
                                                        matrix is a learned BIAS: it is copied as it is and this model
                                                        evaluates it as the bias by default (NeRFNetwork(color_in_pad=1.0),
                                                        DESIGN.md section 2; recalled from upstream, not verifiable here)
    aabb_train, aabb_infer, density_grid, density_bitfield, step_counter      buffers of nerf/renderer.py:32-45

``NeRFNetwork`` here registers the same names with the same shapes and the same flat order (field.py), so the
conversion is a checked copy: prefixes added by DDP / accelerate / torch.compile are stripped, half-precision
parameters are widened to fp32, every shape is verified against this model's level table before anything is written,
and the plain-attribute state of the renderer (``mean_density``, ``iter_density``, ``mean_count``, ``local_step``,
nerf/renderer.py:41-48) is restored when the pickled module carries it.  tiny-cuda-nn is not importable here and the
reference pins no version of it, so the parameter ORDER inside ``params`` is this repo's frozen reading of its
published layout ("parity unpinned" for the external dependency); the shapes are checked, the order cannot be.

A pickled module references classes of ``nerf.*`` and ``tinycudann.*`` that do not exist in this process.
``load_reference_checkpoint`` therefore unpickles with a resolver that substitutes inert stand-ins for every class
outside torch / numpy / the standard containers and then walks ``_parameters`` / ``_buffers`` / ``_modules`` of the
result; no code of the pickled classes is executed.
"""
import collections
import collections.abc
import io
import pickle
import types

import torch

_PREFIXES = ("module.", "_orig_mod.", "nerf.")
_PLAIN_STATE = ("mean_density", "iter_density", "mean_count", "local_step")
# What a pickled module / state_dict legitimately references: tensor rebuilders, storages, dtypes, the containers.  An
# explicit (module, name) allowlist -- NOT whole packages: `builtins`, `torch` or `numpy` as prefixes would let a crafted
# file resolve builtins.eval / getattr / __import__ or torch.hub and REDUCE them.  Everything else becomes an inert stub.
_SAFE_GLOBALS = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"),
    ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "slice"), ("builtins", "list"), ("builtins", "dict"),
    ("builtins", "tuple"), ("builtins", "int"), ("builtins", "float"), ("builtins", "bool"), ("builtins", "complex"),
    ("builtins", "str"), ("builtins", "bytes"), ("builtins", "bytearray"),
    ("_codecs", "encode"), ("copyreg", "_reconstructor"), ("builtins", "object"),
    ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_parameter_with_state"), ("torch._utils", "_rebuild_qtensor"),
    ("torch._tensor", "_rebuild_from_type_v2"), ("torch", "Size"), ("torch", "device"), ("torch", "dtype"),
    ("torch.serialization", "_get_layout"), ("torch.nn.parameter", "Parameter"), ("torch", "Tensor"),
    ("torch.storage", "_load_from_bytes"), ("torch.storage", "UntypedStorage"), ("torch.storage", "TypedStorage"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("numpy", "ndarray"), ("numpy", "dtype"),
}
_SAFE_TORCH_NAMES = {n for n in dir(torch) if n.endswith("Storage")} | {
    "float32", "float16", "bfloat16", "float64", "int32", "int64", "int16", "int8", "uint8", "bool", "float", "half",
    "double", "long", "int", "short"}


class CheckpointError(RuntimeError):
    pass


# --------------------------------------------------------------------------------------------- unpickling without code

class _Stub:
    """Inert stand-in for a class this process cannot import: keeps whatever state the pickle hands it."""

    def __init__(self, *args, **kwargs):
        self._stub_args = args

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self._stub_state = state


def _stub_class(module, name):
    return type(name, (_Stub,), {"__module__": module, "_stub_for": f"{module}.{name}"})


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _SAFE_GLOBALS or (module == "torch" and name in _SAFE_TORCH_NAMES):
            return super().find_class(module, name)
        return _stub_class(module, name)


def _pickle_module():
    mod = types.ModuleType("snerf_stub_pickle")
    mod.Unpickler = _Unpickler
    mod.load = lambda f, **kw: _Unpickler(f, **kw).load()
    mod.loads = lambda b, **kw: _Unpickler(io.BytesIO(b), **kw).load()
    mod.__name__ = "pickle"
    for k in ("PickleError", "UnpicklingError", "PicklingError", "HIGHEST_PROTOCOL", "DEFAULT_PROTOCOL", "Pickler",
              "dump", "dumps"):
        setattr(mod, k, getattr(pickle, k))
    return mod


def _walk(obj, prefix, tensors, plain):
    d = getattr(obj, "__dict__", {})
    for kind in ("_parameters", "_buffers"):
        for k, v in (d.get(kind) or {}).items():
            if isinstance(v, torch.Tensor):
                tensors[prefix + k] = v.detach()
    if not prefix:
        for k in _PLAIN_STATE:
            if k in d:
                plain[k] = d[k]
    for k, m in (d.get("_modules") or {}).items():
        if m is not None:
            _walk(m, prefix + k + ".", tensors, plain)


def extract_state(obj):
    """(tensors, plain) from a state_dict, a {'model': state_dict}-style wrapper, an ``nn.Module`` or an unpickled
    stand-in tree.  tensors: flat name -> tensor; plain: the renderer's non-buffer state when present."""
    if isinstance(obj, collections.abc.Mapping):
        for k in ("state_dict", "model", "nerf"):
            if k in obj and isinstance(obj[k], collections.abc.Mapping) and not isinstance(obj[k], torch.Tensor):
                return extract_state(obj[k])
        tensors = {k: v.detach() for k, v in obj.items() if isinstance(v, torch.Tensor)}
        plain = {k: obj[k] for k in _PLAIN_STATE if k in obj and not isinstance(obj[k], torch.Tensor)}
        return tensors, plain
    tensors, plain = {}, {}
    _walk(obj, "", tensors, plain)
    if not tensors:
        raise CheckpointError(f"no tensors found in checkpoint object of type {type(obj).__name__}")
    return tensors, plain


def _strip(name):
    changed = True
    while changed:
        changed = False
        for p in _PREFIXES:
            if name.startswith(p):
                name, changed = name[len(p):], True
    return name


# ------------------------------------------------------------------------------------------------------------- layout

def reference_state_layout(model):
    """name -> (shape, dtype) of the entries a reference checkpoint of this architecture holds."""
    return collections.OrderedDict((k, (tuple(v.shape), v.dtype)) for k, v in model.state_dict().items())


def describe_params(model):
    """Segments of the two flat ``params`` tensors: [(tensor name, segment, offset, shape)] in storage order."""
    segs = []
    off = 0
    for i, (o, n) in enumerate(model.sigma_net.shapes):
        segs.append(("sigma_net.params", f"mlp.{i}", off, (o, n)))
        off += o * n
    g = model.fdesc.grid
    for l in range(g.n_levels):
        segs.append(("sigma_net.params", f"grid.level{l}" + (".hashed" if g.hashed[l] else ".dense"),
                     off + g.offset[l] * g.n_features, (g.size[l], g.n_features)))
    off = 0
    for i, (o, n) in enumerate(model.color_net.shapes):
        segs.append(("color_net.params", f"mlp.{i}", off, (o, n)))
        off += o * n
    return segs


def load_reference_state_dict(model, state, strict=True, plain=None):
    """Copy a reference state (see module docstring) into ``model``.  Returns (missing, unexpected) name lists; with
    ``strict`` either being non-empty, or any shape mismatch, raises CheckpointError before the model is touched."""
    layout = reference_state_layout(model)
    src = {}
    for k, v in state.items():
        if isinstance(v, torch.Tensor):
            src[_strip(k)] = v
    missing = [k for k in layout if k not in src]
    unexpected = [k for k in src if k not in layout]
    problems = []
    for k, (shape, dtype) in layout.items():
        if k not in src:
            continue
        v = src[k]
        if tuple(v.shape) != shape:
            hint = ""
            if k.endswith(".params"):
                hint = " (config mismatch: check n_levels / log2_hashmap_size / n_neurons / n_hidden_layers / channel_dim)"
            problems.append(f"{k}: checkpoint {tuple(v.shape)} vs model {shape}{hint}")
        elif dtype.is_floating_point != v.dtype.is_floating_point:
            problems.append(f"{k}: checkpoint dtype {v.dtype} vs model {dtype}")
    if strict and (missing or unexpected):
        problems.append(f"missing {missing}, unexpected {unexpected}")
    if problems:
        raise CheckpointError("reference checkpoint does not fit this model: " + "; ".join(problems))
    own = model.state_dict()
    with torch.no_grad():
        for k in layout:
            if k in src:
                own[k].copy_(src[k].to(device=own[k].device, dtype=own[k].dtype))
    for k, v in (plain or {}).items():
        if k in _PLAIN_STATE and hasattr(model, k):
            setattr(model, k, v.item() if isinstance(v, torch.Tensor) else v)
    return missing, unexpected


def load_reference_checkpoint(model, path, strict=True, map_location="cpu"):
    """Load ``nerf.pth`` as written by the reference (pickled module, train.py:307) or a saved state_dict."""
    try:
        obj = torch.load(path, map_location=map_location, pickle_module=_pickle_module(), weights_only=False)
    except CheckpointError:
        raise
    except Exception as e:  # a truncated file or a format torch cannot read: say which file
        raise CheckpointError(f"cannot read checkpoint {path}: {type(e).__name__}: {e}") from e
    tensors, plain = extract_state(obj)
    return load_reference_state_dict(model, tensors, strict=strict, plain=plain)


def reference_state_dict(model):
    """The model's state under the reference's names (fp32, CPU), ready for ``torch.save``; loads into the reference's
    module with ``load_state_dict`` given the same config."""
    return collections.OrderedDict((k, v.detach().to("cpu").clone()) for k, v in model.state_dict().items())


def save_reference_checkpoint(model, path):
    torch.save(reference_state_dict(model), path)

