"""Feature front-end with the interface of honk2's ``AudioProcessor``
(/root/reference/utils/audio_processor.py:7-35), computed by the fused sm_100a kernel in
csrc/mfcc.cu through the C ABI (``kws_mfcc_forward``).

``compute_mfccs`` keeps the reference contract -- one 1-D float waveform in, float32
``(1 + N // hop, n_mels, 1)`` out, values ``2 * ln(mel power)`` with exact zeros left at 0
(audio_processor.py:27-29; the "DCT" there runs over a length-1 axis, i.e. multiplies by 2).
``compute_mfccs_batch`` is the batched form of the collate loop
(/root/reference/data_loader/audio_data_loader.py:26-29): ``[B, N]`` waveforms already on the
GPU -> ``[B, T, n_mels]`` features on the GPU.
"""
import ctypes as C

import numpy as np
import torch

from . import _native


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class AudioProcessor(object):
    def __init__(self, sr=16000, n_dct_filters=40, n_mels=40, f_max=4000, f_min=20, n_fft=480, hop_ms=10):
        super().__init__()
        self.n_mels = n_mels
        self.sr = sr
        self.f_max = f_max if f_max is not None else sr // 2
        self.f_min = f_min
        self.n_fft = n_fft  # 30 ms window size
        self.hop_length = sr // 1000 * hop_ms
        # n_dct_filters is accepted and ignored, as in the reference (audio_processor.py:8).
        self._handles = {}  # device index -> kws_frontend_t*

    # ---- native handle -------------------------------------------------------------------
    def _frontend(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            lib = _native.load()
            out = C.c_void_p()
            with torch.cuda.device(idx):
                _native.check(lib.kws_frontend_create(int(self.sr), int(self.n_mels), float(self.f_min),
                                                      float(self.f_max), int(self.n_fft), int(self.hop_length),
                                                      C.byref(out)), "kws_frontend_create")
            h = self._handles[idx] = out
        return h

    def __del__(self):
        try:
            lib = _native.loaded()   # (never dlopen from a destructor)
            for h in (self._handles.values() if lib is not None else ()):
                lib.kws_frontend_destroy(h)
        except Exception:
            pass

    def __getstate__(self):  # handles never cross a process boundary (DataLoader workers)
        d = dict(self.__dict__)
        d["_handles"] = {}
        return d

    def n_frames(self, n_samples):
        return 1 + n_samples // self.hop_length

    # ---- batched GPU API -----------------------------------------------------------------
    def compute_mfccs_batch(self, waves, out=None):
        """waves: CUDA float32 [B, N] (contiguous) -> CUDA float32 [B, T, n_mels].

        int16 waveforms are taken as 16-bit PCM, the samples of the wav files behind the reference's datasets
        (librosa.core.load returns float32(s / 32768) for them, dataset/gsc_dataset.py:169, hey_snips_dataset.py:74): the kernel converts while
        it stages, the features are bit-identical to those of the converted floats, and the waveforms cost half the
        host->device and HBM bytes."""
        if not (isinstance(waves, torch.Tensor) and waves.is_cuda):
            raise _native.NativeError("compute_mfccs_batch needs a CUDA tensor: there is no CPU path")
        if waves.dim() != 2:
            raise ValueError("compute_mfccs_batch expects waveforms shaped [B, N]")
        pcm16 = waves.dtype == torch.int16
        if not pcm16 and waves.dtype != torch.float32:
            waves = waves.float()
        waves = waves.contiguous()
        B, N = waves.shape
        T = self.n_frames(N)
        if out is None:
            out = torch.empty((B, T, self.n_mels), dtype=torch.float32, device=waves.device)
        elif out.shape != (B, T, self.n_mels) or out.dtype != torch.float32 or not out.is_contiguous() \
                or out.device != waves.device:
            raise ValueError("out must be a contiguous float32 [B, T, n_mels] tensor on the input's device")
        lib = _native.load()
        fe = self._frontend(waves.device)
        with torch.cuda.device(waves.device):
            fn = lib.kws_mfcc_forward_pcm16 if pcm16 else lib.kws_mfcc_forward
            _native.check(fn(fe, C.c_void_p(waves.data_ptr()), B, N, C.c_void_p(out.data_ptr()),
                             _stream_ptr(waves.device)), "kws_mfcc_forward_pcm16" if pcm16 else "kws_mfcc_forward")
        return out

    # ---- streaming windows -------------------------------------------------------------
    @staticmethod
    def n_stream_windows(n_stream_samples, window_size, shift_size):
        """Number of windows a stream yields: StreamingDataset.num_samples (dataset/dataset_utils.py:31)."""
        return max(0, int((n_stream_samples - window_size) / shift_size))

    def compute_mfccs_stream(self, stream, window_size=16000, shift_size=160, first=0, count=None, out=None):
        """Features of the sliding windows of one audio stream, without materialising the windows.

        stream: CUDA float32 [L] (or int16 PCM, converted while staging: bit-identical features); window k =
        stream[k*shift_size : k*shift_size + window_size], exactly what
        StreamingDataset.__getitem__ hands to collate_fn (dataset/dataset_utils.py:72, audio_data_loader.py:26-29;
        gsc_dev_config.json:62-63: 16000 / 160 samples).  Returns CUDA float32 [count, T, n_mels] for windows
        first .. first+count-1 (default: all StreamingDataset.num_samples of them), bit-identical to
        compute_mfccs_batch on the stacked windows.  With shift_size a multiple of the 160-sample hop, 97 of a 1 s
        window's 101 frames are shared with its neighbours and computed once."""
        if not (isinstance(stream, torch.Tensor) and stream.is_cuda):
            raise _native.NativeError("compute_mfccs_stream needs a CUDA tensor: there is no CPU path")
        if stream.dim() != 1:
            raise ValueError("compute_mfccs_stream expects a 1-D stream")
        if window_size < 1 or shift_size < 1:
            raise ValueError("window_size and shift_size must be positive")
        pcm16 = stream.dtype == torch.int16
        if not pcm16 and stream.dtype != torch.float32:
            stream = stream.float()
        stream = stream.contiguous()
        total = self.n_stream_windows(stream.numel(), window_size, shift_size)
        if count is None:
            count = total - first
        if first < 0 or count < 0 or first + count > total:
            raise ValueError(f"windows {first}..{first + count} outside the stream's {total} windows")
        T = self.n_frames(window_size)
        if out is None:
            out = torch.empty((count, T, self.n_mels), dtype=torch.float32, device=stream.device)
        elif out.shape != (count, T, self.n_mels) or out.dtype != torch.float32 or not out.is_contiguous() \
                or out.device != stream.device:
            raise ValueError("out must be a contiguous float32 [count, T, n_mels] tensor on the input's device")
        if count == 0:
            return out
        lib = _native.load()
        fe = self._frontend(stream.device)
        need = lib.kws_mfcc_stream_scratch_bytes(fe, count, window_size, shift_size)
        scratch = torch.empty(max(need, 16), dtype=torch.uint8, device=stream.device)
        base = stream.data_ptr() + stream.element_size() * first * shift_size
        fn = lib.kws_mfcc_stream_forward_pcm16 if pcm16 else lib.kws_mfcc_stream_forward
        with torch.cuda.device(stream.device):
            _native.check(fn(fe, C.c_void_p(base), count, window_size, shift_size, C.c_void_p(out.data_ptr()),
                             C.c_void_p(scratch.data_ptr()), need, _stream_ptr(stream.device)),
                          "kws_mfcc_stream_forward_pcm16" if pcm16 else "kws_mfcc_stream_forward")
        return out

    # ---- reference API -------------------------------------------------------------------
    def compute_mfccs(self, data, device=None):
        """1-D float np.ndarray -> float32 np.ndarray (T, n_mels, 1), like the reference.
        Input checks mirror librosa.util.valid_audio, which the reference call goes through."""
        if not isinstance(data, np.ndarray):
            raise ValueError("Audio data must be of type numpy.ndarray")
        if not np.issubdtype(data.dtype, np.floating):
            raise ValueError("Audio data must be floating-point")
        if data.ndim != 1:
            raise ValueError(f"Invalid shape for monophonic audio: ndim={data.ndim}, shape={data.shape}")
        if not np.isfinite(data).all():
            raise ValueError("Audio buffer is not finite everywhere")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        wave = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).to(dev).unsqueeze(0)
        feat = self.compute_mfccs_batch(wave)
        return feat[0].unsqueeze(-1).cpu().numpy()

    def compute_pcen(self, data):
        raise NotImplementedError(
            "PCEN is outside the hot path: the reference depends on the un-vendored pytorch-pcen package "
            "and no shipped config selects it (audio_data_loader.py:30-33)")
