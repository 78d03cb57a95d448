"""Mirror of lib/utils/preprocess_sample.py:7-56 for the test-time image path (Faster / Mask R-CNN: no proposals in the sample; Fast R-CNN:
pre-computed proposals, single-level or distributed over the FPN levels): sample['image'] comes back as a CUDA FloatTensor [1,3,Hp,Wp]
built on the device."""
import numpy as np
import torch

from .blob import image_to_blob
from .multilevel_rois import add_multilevel_rois_for_test


class preprocess_sample(object):
    def __init__(self, target_sizes=800, max_size=1333, mean=[122.7717, 115.9465, 102.9801], remove_dup_proposals=True, fpn_on=False,
                 spatial_scale=0.0625, sample_proposals_for_training=False):
        self.mean = mean
        self.target_sizes = target_sizes if isinstance(target_sizes, list) else [target_sizes]
        self.max_size = max_size
        self.remove_dup_proposals = remove_dup_proposals
        self.fpn_on = fpn_on
        self.spatial_scale = spatial_scale
        if sample_proposals_for_training:
            raise NotImplementedError("training-time proposal sampling is outside the inference path of this build")

    def __call__(self, sample):
        original_im_size = sample['image'].shape
        blob, scale = image_to_blob(sample['image'], self.mean, self.target_sizes[0], self.max_size, self.fpn_on)
        sample['image'] = blob
        sample['scaling_factors'] = scale
        sample['original_im_size'] = torch.FloatTensor(original_im_size)
        if 'dbentry' in sample:
            boxes = sample['dbentry']['boxes']
            if len(boxes) != 0:                                   # Fast R-CNN test: pre-computed proposals (preprocess_sample.py:37-47)
                proposals = boxes * scale
                if self.remove_dup_proposals:
                    proposals, _ = self.remove_dup_prop(proposals)
                if not self.fpn_on:
                    sample['rois'] = torch.FloatTensor(proposals)
                else:                                              # rois, rois_fpn2..5, rois_idx_restore_int32 (preprocess_sample.py:43-46)
                    for k, v in add_multilevel_rois_for_test({'rois': proposals}, 'rois').items():
                        sample[k] = torch.FloatTensor(v)
            del sample['dbentry']
        return sample

    def remove_dup_prop(self, proposals):                         # preprocess_sample.py:61-69
        v = np.array([1e3, 1e6, 1e9, 1e12])
        hashes = np.round(proposals * self.spatial_scale).dot(v)
        _, index, inv_index = np.unique(hashes, return_index=True, return_inverse=True)
        return proposals[index, :], inv_index
