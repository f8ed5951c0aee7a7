"""Drop-in for the reference's ExternalInfoLoaders/InstanceSegmentationLoader.py
(/root/reference/ExternalInfoLoaders/InstanceSegmentationLoader.py:12-121): the instance label map and the
per-instance areas cross PCIe as int64, the clamp + GloVe gather (+ human sizes) happen on the GPU.
"""
import sys

import numpy as np
import torch

from .. import ops
from .SemanticsLoader import _data_file


class InstanceSegmentationLoader():
    def __init__(self, args, device=None):
        self.args = args
        self.device = torch.device("cuda") if device is None else torch.device(device)
        self.embeddings_path = None
        self.human_sizes_path = None
        self.word_embeddings_semantics = None
        self.background_class_num = None
        self.human_sizes = None
        self._dev_tables = {}
        self.clamp_host_batch = False  # see SemanticsLoader: the clamp happens on the device copy
        self.set_embeddings_path()
        self.set_human_sizes_path()
        self.load_word_embeddings()
        self.load_human_sizes()

    def set_embeddings_path(self):
        mode = self.args.use_instance_segmentation
        if mode is None:
            return
        if mode == "raw":
            sys.exit("Error: raw instance semantics not implemented")
        if mode == "coco":
            self.embeddings_path = _data_file("coco_81_classes_maskrcnn_ordering_glove_twitter_27b_25d_embeddings.npy")
            self.background_class_num = 0
        elif "ade20k_swin" in mode:
            self.embeddings_path = _data_file("ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy")
            self.background_class_num = 100
        assert self.embeddings_path is not None
        assert self.background_class_num is not None

    def set_human_sizes_path(self):
        mode = self.args.use_instance_segmentation
        if mode is not None and "ade20k_swin" in mode and "human_sizes" in mode:
            self.human_sizes_path = _data_file(
                "ade20k_classes_abs_sizes_shuffled.npy" if "shuffled" in mode else "ade20k_classes_abs_sizes.npy")

    def load_word_embeddings(self):
        if self.embeddings_path is not None:
            self.word_embeddings_semantics = torch.from_numpy(np.load(self.embeddings_path))

    def load_human_sizes(self):
        if self.human_sizes_path is not None:
            self.human_sizes = torch.from_numpy(np.load(self.human_sizes_path))

    def _table(self, name, host, dtype):
        key = (name, dtype)
        if key not in self._dev_tables:
            self._dev_tables[key] = host.to(dtype).contiguous().to(self.device)
        return self._dev_tables[key]

    def get_instance_segmentation(self, batch):
        """-> (labels_raw clamped [B,1,H,W] int64, embedding float64 [B,25,H,W], areas float32 [B,1|4,H,W])."""
        if self.word_embeddings_semantics is None:
            return None, None, None
        host_raw = batch['instance_labels']
        raw = host_raw.to(self.device, non_blocking=True).contiguous()
        areas_raw = batch['instance_areas'].to(self.device, non_blocking=True)
        bg = self.background_class_num
        # the embedding stays float64 in the reference (no .float() at :107-109)
        table = self._table("emb", self.word_embeddings_semantics, torch.float64)
        if raw.dtype == torch.int64:
            emb = ops.gather_embed(raw, table, background=bg, write_back=True)
        else:  # int32 instance maps as stored on disk (.npz 'arr_0', dataloader.py:136-150): 4 B/px over the bus
            raw64 = torch.empty(raw.shape, dtype=torch.int64, device=raw.device)
            emb = ops.gather_embed(raw, table, background=bg, labels_out=raw64)
            raw = raw64
        areas = ops.cast_i64_f32(areas_raw) if areas_raw.dtype == torch.int64 else areas_raw.float()
        if self.human_sizes is not None:
            sizes = ops.gather_embed(raw, self._table("sizes", self.human_sizes, torch.float32), background=None)
            areas = torch.cat((areas, sizes), dim=1)
        if self.clamp_host_batch and isinstance(host_raw, torch.Tensor) and not host_raw.is_cuda:
            rows = self.word_embeddings_semantics.shape[0]
            host_raw[host_raw < 0] = bg
            host_raw[host_raw > rows - 1] = bg
        return raw, emb, areas
