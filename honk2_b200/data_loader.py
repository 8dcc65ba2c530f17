"""``data_loader.AudioDataLoader`` with the reference's constructor and iteration contract
(/root/reference/data_loader/audio_data_loader.py:10-35): ``AudioDataLoader(data_loader_config, dataset)`` iterates
``(FloatTensor[B, T, 40], LongTensor[B])`` batches.

The reference's ``collate_fn`` runs ``compute_mfccs`` once per sample inside the DataLoader workers and grows the
batch with ``torch.cat`` (:26-29).  Here the workers (which must not touch CUDA: they are forked) only stack the raw
waveforms; the main process moves the ``[B, N]`` batch to the GPU once and runs the fused front-end kernel on all of
it (``AudioProcessor.compute_mfccs_batch``).  The features come back as CUDA tensors, so ``data.to(device)`` in
``evaluate`` (run/test.py:23) is a no-op.
"""
import numpy as np
import torch
from torch.utils.data import DataLoader

from .audio_processor import AudioProcessor
from .class_registry import register_cls


@register_cls('data_loader.AudioDataLoader')
class AudioDataLoader(DataLoader):
    def __init__(self, data_loader_config, dataset, device=None):
        self.audio_preprocessing = data_loader_config["audio_preprocessing"]
        if self.audio_preprocessing != "MFCCs":
            # the reference's PCEN branch is broken (audio_data_loader.py:30-33 uses an unimported numpy) and depends on
            # the un-vendored pytorch-pcen package; no shipped config selects it
            raise NotImplementedError(f"audio_preprocessing {self.audio_preprocessing!r}: only 'MFCCs' is on the hot path")
        self.audio_processor = AudioProcessor()
        self.device = device
        super().__init__(
            dataset=dataset,
            batch_size=data_loader_config["batch_size"],
            shuffle=data_loader_config["shuffle"],
            collate_fn=self.collate_fn,
            # (the model-zoo configs carry no num_workers key and make the reference raise KeyError, :19-21)
            num_workers=data_loader_config.get("num_workers", 0),
            pin_memory=torch.cuda.is_available())

    @staticmethod
    def collate_fn(batch):
        """Worker side: raw waveforms stacked as float32 [B, N] (or a list when the lengths differ) + targets."""
        waves = [np.asarray(sample, dtype=np.float32) for sample, _ in batch]
        targets = torch.tensor([label for _, label in batch])
        if len({w.shape[0] for w in waves}) == 1:
            return torch.from_numpy(np.stack(waves)), targets
        return [torch.from_numpy(w) for w in waves], targets

    def features(self, waves):
        """[B, N] waveforms (CPU or CUDA) -> CUDA float32 [B, T, n_mels]."""
        dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        if isinstance(waves, (list, tuple)):   # ragged batch: one call per clip, like the reference's loop
            return torch.cat([self.audio_processor.compute_mfccs_batch(w.to(dev).unsqueeze(0)) for w in waves], 0)
        return self.audio_processor.compute_mfccs_batch(waves.to(dev, non_blocking=True))

    def __iter__(self):
        for waves, targets in super().__iter__():
            yield self.features(waves), targets
