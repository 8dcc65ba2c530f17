"""Host side of the streaming-window path: what ``StreamingDataset`` (dataset/dataset_utils.py:20-95)
does around the audio, restated for a stream that is already one array -- window count, window targets -- plus
``evaluate_stream``: features of all sliding windows from the shared-frame front-end, then the model, batch by batch.
"""
import numpy as np
import torch

from .audio_processor import AudioProcessor


def n_stream_windows(total_num_samples, window_size, shift_size):
    """StreamingDataset.num_samples (dataset_utils.py:31)."""
    return AudioProcessor.n_stream_windows(total_num_samples, window_size, shift_size)


def stream_window_targets(segment_lengths, segment_labels, n_labels, window_size, shift_size, n_windows=None):
    """Target label of every window of a stream made of labelled segments (audio files played back to back).

    The reference keeps a running ``label_counter`` over the samples of the current window and picks the label with
    the largest count, the LOWEST label index winning ties (strict ``>`` while enumerating, dataset_utils.py:74-81).
    Its incremental bookkeeping (:58-62, :84-92) always equals the label histogram of
    stream[k*shift : k*shift + window]; this computes all windows at once from per-label prefix sums.
    Returns int64 [n_windows]."""
    segment_lengths = np.asarray(segment_lengths, dtype=np.int64)
    segment_labels = np.asarray(segment_labels, dtype=np.int64)
    if segment_lengths.shape != segment_labels.shape or segment_lengths.ndim != 1:
        raise ValueError("segment_lengths and segment_labels must be 1-D and of equal length")
    if (segment_lengths < 0).any() or ((segment_labels < 0) | (segment_labels >= n_labels)).any():
        raise ValueError("negative segment length or label outside [0, n_labels)")
    total = int(segment_lengths.sum())
    avail = n_stream_windows(total, window_size, shift_size)
    if n_windows is None:
        n_windows = avail
    if n_windows > avail:
        raise ValueError(f"{n_windows} windows requested, the stream has {avail}")
    if n_windows <= 0:
        return np.zeros((0,), dtype=np.int64)
    # prefix[c, i] = number of samples with label c among stream[:i], evaluated only where windows start / end
    bounds = np.concatenate(([0], np.cumsum(segment_lengths)))
    starts = np.arange(n_windows, dtype=np.int64) * shift_size
    ends = starts + window_size
    counts = np.zeros((n_windows, n_labels), dtype=np.int64)
    for c in range(n_labels):
        seg_c = np.where(segment_labels == c, segment_lengths, 0)
        cum_c = np.concatenate(([0], np.cumsum(seg_c)))          # label-c samples in the first j segments

        def prefix(pos):
            j = np.searchsorted(bounds, pos, side="right") - 1    # segment containing position pos
            j = np.minimum(j, len(segment_lengths) - 1)
            inside = pos - bounds[j]
            return cum_c[j] + np.where(segment_labels[j] == c, inside, 0)

        counts[:, c] = prefix(ends) - prefix(starts)
    return np.argmax(counts, axis=1).astype(np.int64)


def evaluate_stream(model, audio_processor, stream, window_size=16000, shift_size=160, targets=None,
                    batch_size=8192):
    """Logits (and accuracy, if targets are given) of every window of a CUDA stream tensor [L].

    Replaces iterating a StreamingDataset through AudioDataLoader + run/test.py:evaluate (18-41): per batch ONE
    front-end call that computes each shared frame once (AudioProcessor.compute_mfccs_stream) and one model call;
    correct/total accumulate on the device, one host sync at the end."""
    if not (isinstance(stream, torch.Tensor) and stream.is_cuda and stream.dim() == 1):
        raise ValueError("evaluate_stream expects a 1-D CUDA tensor")
    n = n_stream_windows(stream.numel(), window_size, shift_size)
    from .metric import Acc
    logits = []
    acc = Acc()
    tgt = None
    if targets is not None:
        tgt = torch.as_tensor(np.asarray(targets), dtype=torch.int64, device=stream.device)
        if tgt.numel() != n:
            raise ValueError(f"{tgt.numel()} targets for {n} windows")
    with torch.no_grad():
        for first in range(0, n, batch_size):
            count = min(batch_size, n - first)
            feats = audio_processor.compute_mfccs_stream(stream, window_size, shift_size, first=first, count=count)
            y = model(feats)
            logits.append(y)
            if tgt is not None:
                acc.accumulate(y, tgt[first:first + count])     # kws_acc_accumulate: counts stay on the device
    out = torch.cat(logits) if logits else torch.empty((0, 0), device=stream.device)
    if tgt is None:
        return out
    return out, {"correct": acc.counts()[0], "total": n}
