"""CPU ORACLE (test infrastructure, NOT product code) -- restatement of the window / target bookkeeping of
``StreamingDataset.__getitem__`` (/root/reference/dataset/dataset_utils.py:47-95) for a stream given as labelled
segments (``_load_sample`` returns one audio file and its label, :39-41).  Pure-Python loops: small cases only.

PARITY PIN: the reference holds no tests or fixtures for this path; the restatement follows the reference statement by
statement (lazy loading :52-66, window :69, strict-``>`` majority vote :72-79, counter update :82-90, drop :92-93).

Only tests/ may import this module.
"""
import numpy as np


def iterate_windows(segments, labels, n_labels, window_size, shift_size, num_samples):
    """segments: list of 1-D arrays; yields (audio_window, target_label) for index 0 .. num_samples-1."""
    loaded_data = np.array([])
    loaded_labels = []
    audio_file_idx = 0
    label_counter = n_labels * [0]
    for _ in range(num_samples):
        while len(loaded_labels) < window_size:
            audio_data, label = segments[audio_file_idx], labels[audio_file_idx]
            loaded_data = np.concatenate((loaded_data, audio_data), axis=0)
            prev = len(loaded_labels)
            loaded_labels += len(audio_data) * [label]
            if prev < window_size:
                label_counter[label] += min(window_size, len(loaded_data)) - prev
            audio_file_idx += 1
        audio_window = loaded_data[:window_size]
        max_count, target = 0, None
        for label, count in enumerate(label_counter):
            if count > max_count:
                target, max_count = label, count
        for i in range(shift_size):
            label_counter[loaded_labels[i]] -= 1
        for i in range(shift_size):
            idx = window_size + i
            if len(loaded_labels) > idx:
                label_counter[loaded_labels[idx]] += 1
        loaded_labels = loaded_labels[shift_size:]
        loaded_data = loaded_data[shift_size:]
        yield audio_window, target
