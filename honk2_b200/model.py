"""``model.ResNet`` / ``model.CNN`` with honk2's construction and call contract, executed by the
native sm_100a library.

Each class builds the SAME ``self.layers`` ModuleDict as the reference constructor
(/root/reference/model/resnet.py:11-36, /root/reference/model/cnn.py:12-77), in the same order,
so ``state_dict()`` keys, ``load_state_dict`` (utils/workspace.py:61), default initialisation
under ``torch.manual_seed`` and ``print(model)`` are identical.  ``forward(x)`` keeps the
reference signature -- ``x`` float32 ``[B, T, F]`` -> raw logits ``[B, n_labels]``
(resnet.py:38-60, cnn.py:79-107) -- but runs eval-mode inference through
``kws_model_forward``; the torch submodules only hold the parameters.  There is no CPU or
eager fallback: a CPU tensor, training mode, or a missing library raises.

Optional config key (default keeps old configs working): ``"precision": "fp32" | "bf16" | "bf16x3"``
(fp32 = CUDA-core FFMA reference mode; bf16 = tcgen05 tensor cores with bf16 operands; bf16x3 = tensor cores with
split-bf16 operands -- activations and weights as hi + lo bf16 pairs, three MMAs per product, fp32 accumulate --
which meets the fp32 tolerance at tensor-core speed).
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _native
from .class_registry import register_cls
from .torch_utils import calculate_conv_output_size, calculate_pool_output_size


class BaseModel(nn.Module):
    """/root/reference/model/model_utils.py:6-11 plus the native-handle plumbing."""

    def __init__(self):
        super().__init__()
        self._native_state = {}   # device index -> dict(handle, stamp, ws)
        self.precision = "fp32"
        self.chunk = {"fp32": 0, "bf16": 0}

    def num_params(self):
        return sum(p.numel() for p in self.parameters())

    def num_trainable_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    # ---- to be provided by subclasses ----------------------------------------------------
    def _create_handle(self, lib):
        raise NotImplementedError

    def _upload(self, lib, handle, stream):
        raise NotImplementedError

    # ---- native plumbing -----------------------------------------------------------------
    def _tensors(self):
        return list(self.parameters()) + list(self.buffers())

    def _stamp(self):
        return tuple((t.data_ptr(), t._version) for t in self._tensors())

    def _state(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        st = self._native_state.get(idx)
        lib = _native.load()
        if st is None:
            with torch.cuda.device(idx):
                st = {"handle": self._create_handle(lib), "stamp": None, "ws": None, "chunk": None}
            self._native_state[idx] = st
        for t in self._tensors():
            if t.device != device:
                raise _native.NativeError(
                    f"model tensors live on {t.device} but the input is on {device}; call model.to(device)")
        stamp = self._stamp()
        # A torch.nn.DataParallel replica (run/test.py:69-70) carries fresh broadcast copies of the weights on every call, and
        # the caching allocator may hand them the addresses of the previous call's copies: always repack for a replica.
        if st["stamp"] != stamp or getattr(self, "_is_replica", False):   # first call, load_state_dict, .to(), in-place update
            with torch.cuda.device(idx):
                self._upload(lib, st["handle"], C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
                # the contiguous fp32 staging copies made by _upload die when it returns
                torch.cuda.current_stream(device).synchronize()
            st["stamp"] = stamp
        chunk = tuple(self.chunk.get(name, 0) for name in _native.PRECISIONS)
        if st["chunk"] != chunk:
            for name, prec in _native.PRECISIONS.items():
                _native.check(lib.kws_model_set_chunk(st["handle"], prec, int(self.chunk.get(name, 0))),
                              "kws_model_set_chunk")
            st["chunk"] = chunk
        return lib, st

    def _workspace(self, st, nbytes, device):
        ws = st["ws"]
        if ws is None or ws.numel() < nbytes or ws.device != device:
            ws = st["ws"] = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        return ws

    def _precision_id(self):
        if self.precision not in _native.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_native.PRECISIONS)}, got {self.precision!r}")
        return _native.PRECISIONS[self.precision]

    def _check_input(self, x, dims, pcm16_ok=False):
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise _native.NativeError("honk2_b200 models run on a B200 only: move the input (and the model) "
                                      "to a CUDA device; there is no CPU fallback")
        if self.training:
            raise _native.NativeError("honk2_b200 implements eval-mode inference only: call model.eval() "
                                      "(run/test.py:21)")
        if x.dim() != dims:
            raise ValueError(f"expected a {dims}-D input, got shape {tuple(x.shape)}")
        if x.dtype != torch.float32 and not (pcm16_ok and x.dtype == torch.int16):
            x = x.float()
        return x.contiguous()

    def _logits_out(self, out, B, device):
        if out is None:
            return torch.empty((B, self.n_labels), dtype=torch.float32, device=device)
        if out.shape != (B, self.n_labels) or out.dtype != torch.float32 or not out.is_contiguous() \
                or out.device != device:
            raise ValueError("out must be a contiguous float32 [B, n_labels] tensor on the input's device")
        return out

    def forward(self, x, out=None):
        x = self._check_input(x, 3)
        B, T, F = x.shape
        lib, st = self._state(x.device)
        prec = self._precision_id()
        logits = self._logits_out(out, B, x.device)
        if B == 0:
            return logits
        with torch.cuda.device(x.device):
            need = lib.kws_model_workspace_bytes(st["handle"], B, T, F, prec)
            if need == 0:
                raise _native.NativeError(f"{type(self).__name__}: precision {self.precision!r} is not available "
                                          f"for input {T}x{F}")
            ws = self._workspace(st, need, x.device)
            _native.check(lib.kws_model_forward(st["handle"], C.c_void_p(x.data_ptr()), B, T, F,
                                                C.c_void_p(logits.data_ptr()), prec, C.c_void_p(ws.data_ptr()),
                                                ws.numel(), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                          "kws_model_forward")
        return logits

    def forward_wave(self, waves, audio_processor, out=None):
        """Fused collate + forward: CUDA float32 waveforms [B, N] -> logits [B, n_labels]
        (data_loader/audio_data_loader.py:26-29 followed by model(x)).  `out`: optional preallocated logits.
        int16 waveforms are 16-bit PCM samples (see AudioProcessor.compute_mfccs_batch): same logits, bit for bit, as
        for waves.float() / 32768."""
        waves = self._check_input(waves, 2, pcm16_ok=True)
        pcm16 = waves.dtype == torch.int16
        B, N = waves.shape
        lib, st = self._state(waves.device)
        fe = audio_processor._frontend(waves.device)
        prec = self._precision_id()
        logits = self._logits_out(out, B, waves.device)
        if B == 0:
            return logits
        with torch.cuda.device(waves.device):
            need = lib.kws_model_wave_workspace_bytes(st["handle"], fe, B, N, prec)
            if need == 0:
                raise _native.NativeError(f"{type(self).__name__}: precision {self.precision!r} is not available")
            ws = self._workspace(st, need, waves.device)
            fn = lib.kws_model_forward_wave_pcm16 if pcm16 else lib.kws_model_forward_wave
            _native.check(fn(
                st["handle"], fe, C.c_void_p(waves.data_ptr()), B, N, C.c_void_p(logits.data_ptr()), prec,
                C.c_void_p(ws.data_ptr()), ws.numel(),
                C.c_void_p(torch.cuda.current_stream(waves.device).cuda_stream)),
                "kws_model_forward_wave_pcm16" if pcm16 else "kws_model_forward_wave")
        return logits

    def last_launches(self, device=None):
        """Kernel launches issued by the last forward on `device` (bench 'gpu_launches')."""
        idx = torch.cuda.current_device() if device is None else torch.device(device).index
        st = self._native_state.get(idx)
        return 0 if st is None else int(_native.load().kws_model_last_launches(st["handle"]))

    def __del__(self):
        # (never dlopen from a destructor: a model that was only ever used on the CPU side -- state_dict, parameter
        # counts -- must not map the native library into its process)
        # A torch.nn.DataParallel replica (run/test.py:69-70) is a shallow copy that SHARES this dict of handles with the
        # module it was made from and dies after every forward: only the original owns (and frees) the handles.
        if getattr(self, "_is_replica", False):
            return
        try:
            lib = _native.loaded()
            if lib is not None:
                for st in self._native_state.values():
                    lib.kws_model_destroy(st["handle"])
                self._native_state.clear()
        except Exception:
            pass

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_native_state"] = {}
        return d


@register_cls('model.ResNet')
class ResNet(BaseModel):
    def __init__(self, config):
        super().__init__()
        self.n_layers = config["n_layers"]
        n_maps = config["n_feature_maps"]
        self.n_maps = n_maps
        self.n_labels = config["n_labels"]
        self.use_dilation = bool(config["use_dilation"])
        self.pool = tuple(_pair(config["pool"])) if "pool" in config else None
        self.precision = config.get("precision", "fp32")

        self.layers = nn.ModuleDict()
        self.layers["conv_0"] = nn.Conv2d(1, n_maps, (3, 3), padding=1, bias=False)
        for i in range(1, self.n_layers + 1):
            d = int(2 ** ((i - 1) // 3)) if config["use_dilation"] else 1
            self.layers[f"conv_{i}"] = nn.Conv2d(n_maps, n_maps, (3, 3), padding=d, dilation=d, bias=False)
            self.layers[f"bn_{i}"] = nn.BatchNorm2d(n_maps, affine=False)
        if "pool" in config:
            self.layers["pool"] = nn.AvgPool2d(config["pool"])
        self.layers["output"] = nn.Linear(n_maps, config["n_labels"])
        self.activations = nn.ModuleDict({"relu": nn.ReLU()})

    def _create_handle(self, lib):
        cfg = _native.ResNetConfig(self.n_layers, self.n_maps, int(self.use_dilation),
                                   self.pool[0] if self.pool else 0, self.pool[1] if self.pool else 0,
                                   self.n_labels)
        out = C.c_void_p()
        _native.check(lib.kws_resnet_create(C.byref(cfg), C.byref(out)), "kws_resnet_create")
        return out

    def _upload(self, lib, handle, stream):
        L = self.layers
        n = self.n_layers
        keep = []  # contiguous fp32 views must outlive the call

        def ptr(t):
            t = t.detach().to(torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        arr = C.c_void_p * max(n, 1)
        conv = arr(*[ptr(L[f"conv_{i}"].weight) for i in range(1, n + 1)])
        mean = arr(*[ptr(L[f"bn_{i}"].running_mean) for i in range(1, n + 1)])
        var = arr(*[ptr(L[f"bn_{i}"].running_var) for i in range(1, n + 1)])
        w = _native.ResNetWeights(ptr(L["conv_0"].weight), conv, mean, var, ptr(L["output"].weight),
                                  ptr(L["output"].bias))
        _native.check(lib.kws_resnet_set_weights(handle, C.byref(w), stream), "kws_resnet_set_weights")
        torch.cuda.current_stream().synchronize()  # `keep` must outlive the device-side repack


def _pair(v):
    if isinstance(v, (list, tuple)):
        return int(v[0]), int(v[1])
    return int(v), int(v)


@register_cls('model.CNN')
class CNN(BaseModel):
    def __init__(self, config):
        super().__init__()
        self.layers = nn.ModuleDict()
        self.n_labels = config["n_labels"]
        self.precision = config.get("precision", "fp32")
        self.config = {k: config[k] for k in ("time", "frequency", "conv_0", "pool_0", "conv_1", "pool_1",
                                              "lin_0", "dnn_0", "dnn_1") if k in config}

        # Same submodules, same order and same RNG consumption as the reference constructor (cnn.py:14-73), so that
        # state_dict keys and default initialisation agree: [conv_i, pool_i]* -> lin_0 / dnn_0 / dnn_1 -> lin_1.
        channels, size = 1, [config["time"], config["frequency"]]
        for i in (0, 1):
            if f"conv_{i}" not in config:
                break
            spec = config[f"conv_{i}"]
            self.layers[f"conv_{i}"] = nn.Conv2d(channels, spec["out_channels"], spec["kernel_size"], stride=spec["stride"])
            channels = spec["out_channels"]
            size = calculate_conv_output_size(size, spec["kernel_size"], stride=spec["stride"])
            pool = config[f"pool_{i}"]["kernel_size"]
            self.layers[f"pool_{i}"] = nn.MaxPool2d(pool)
            size = calculate_pool_output_size(size, pool)
        width = int(channels * np.prod(size))
        for name in ("lin_0", "dnn_0", "dnn_1"):
            if name in config:
                self.layers[name] = nn.Linear(width, config[name]["out_features"])
                width = config[name]["out_features"]
        self.layers["lin_1"] = nn.Linear(width, config["n_labels"])
        self.layers["dropout"] = nn.Dropout(config["dropout_prob"])
        self.activations = nn.ModuleDict({"relu": nn.ReLU()})

    def _create_handle(self, lib):
        c = self.config
        k0, s0, p0 = _pair(c["conv_0"]["kernel_size"]), _pair(c["conv_0"]["stride"]), _pair(c["pool_0"]["kernel_size"])
        if "conv_1" in c:
            c1 = (c["conv_1"]["out_channels"], *_pair(c["conv_1"]["kernel_size"]), *_pair(c["conv_1"]["stride"]),
                  *_pair(c["pool_1"]["kernel_size"]))
        else:
            c1 = (0, 0, 0, 0, 0, 0, 0)
        outs = [c[n]["out_features"] if n in c else 0 for n in ("lin_0", "dnn_0", "dnn_1")]
        cfg = _native.CnnConfig(c["time"], c["frequency"], c["conv_0"]["out_channels"], *k0, *s0, *p0, *c1, *outs,
                                self.n_labels)
        out = C.c_void_p()
        _native.check(lib.kws_cnn_create(C.byref(cfg), C.byref(out)), "kws_cnn_create")
        return out

    def _upload(self, lib, handle, stream):
        L = self.layers
        keep = []

        def ptr(name, attr):
            if name not in L:
                return None
            t = getattr(L[name], attr).detach().to(torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        fields = []
        for name in ("conv_0", "conv_1", "lin_0", "dnn_0", "dnn_1", "lin_1"):
            fields += [ptr(name, "weight"), ptr(name, "bias")]
        w = _native.CnnWeights(*fields)
        _native.check(lib.kws_cnn_set_weights(handle, C.byref(w), stream), "kws_cnn_set_weights")
        torch.cuda.current_stream().synchronize()
