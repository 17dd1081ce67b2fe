"""B200-native UNETR hot path of ilkyyldz95/3DmedicalImageSegmentation.

Public surface (names follow the reference):
    UNETR                                  unetr.py:21 (returns (enc4, logits))
    MonaiUNETR                             monai.networks.nets.UNETR as used at unetr_segmentation_3d.py:36
    DiceCELoss                             unetr_segmentation_3d.py:404
    extract_triplets_more_partitions, BTLoss   unetr_ranking_pretraining_3d.py:59,202
    sliding_window_inference               unetr_segmentation_3d.py:109,143,694
    DiceMetric, ConfusionMatrixMetric      monai.metrics as constructed at unetr_segmentation_3d.py:485-494 (validation tail)
    FusedAdamW                             torch.optim.AdamW at unetr_segmentation_3d.py:522 / unetr_ranking_pretraining_3d.py:466
    GraphedTrainStep                       the training-loop body of unetr_segmentation_3d.py:218-226 replayed as one CUDA graph
    transforms.{Compose, RandCropByPosNegLabeld, RandFlipd, RandRotate90d, RandShiftIntensityd, RandSpatialCropSamplesd,
                ConvertToMultiChannelBasedOnBratsClassesd}   GPU-side tail of the training transforms, seg:65-93,341-375 / rank:365-369
All arithmetic runs in csrc/libunetr_b200.so (hand-written sm_100a CUDA behind include/unetr_b200.h).
"""
from . import _lib
from . import transforms
from .inferers import shard_windows, sliding_window_inference, window_starts
from .losses import (BTLoss, DiceCELoss, configure_ranking, extract_triplets_more_partitions, ranking_loss)
from .metrics import (ConfusionMatrixMetric, DiceMetric, segmentation_counts, segmentation_counts_from_label_maps)
from .graph import GraphedTrainStep
from .optim import FusedAdamW
from .unetr import UNETR, MonaiUNETR

__all__ = ["UNETR", "MonaiUNETR", "DiceCELoss", "BTLoss", "extract_triplets_more_partitions", "ranking_loss",
           "configure_ranking", "sliding_window_inference", "window_starts", "shard_windows", "FusedAdamW", "GraphedTrainStep", "transforms",
           "DiceMetric", "ConfusionMatrixMetric", "segmentation_counts", "segmentation_counts_from_label_maps"]
