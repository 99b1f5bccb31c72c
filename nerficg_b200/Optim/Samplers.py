"""Ray-batch samplers of the trainer (reference src/Optim/Samplers/{DatasetSamplers,ImageSamplers,utils}.py)."""
from __future__ import annotations

import torch

from .. import Framework


class RandomImageSampler:
    """Uniform random pixel ids with replacement (ImageSamplers.py:42-45); drawn on the CPU generator like the reference."""

    def __init__(self, num_elements: int) -> None:
        self.num_elements = num_elements

    def get(self, ray_batch_size: int) -> torch.Tensor:
        return torch.randint(low=0, high=self.num_elements, size=(ray_batch_size,))


class RandomSequentialSampler:
    """Random permutation walked sequentially, reshuffled when exhausted."""

    def __init__(self, num_elements: int) -> None:
        self.num_elements = num_elements
        self.order = torch.randperm(num_elements)
        self.cursor = 0

    def get(self, num_samples: int = 1) -> torch.Tensor:
        out = []
        for _ in range(num_samples):
            if self.cursor >= self.num_elements:
                self.order = torch.randperm(self.num_elements)
                self.cursor = 0
            out.append(self.order[self.cursor])
            self.cursor += 1
        return torch.stack(out)


class SequentialSampler:
    """0, 1, ..., n-1, 0, ... (Samplers/utils.py: the ``random=False`` view order)."""

    def __init__(self, num_elements: int) -> None:
        self.num_elements = num_elements
        self.cursor = 0

    def get(self, num_samples: int = 1) -> torch.Tensor:
        out = (torch.arange(num_samples) + self.cursor) % self.num_elements
        self.cursor = int((self.cursor + num_samples) % self.num_elements)
        return out


def _takes_pixel_ids(view) -> bool:
    """Our View.get_rays(pixel_ids) generates only the sampled rays on the device (K0); the reference's
    View.get_rays() (src/Datasets/utils.py:1053) takes no argument and returns the whole image's RayBatch."""
    import inspect
    try:
        return len(inspect.signature(view.get_rays).parameters) >= 1
    except (TypeError, ValueError):
        return False


class DatasetSampler:
    """One view per call (random permutation or sequential), then pixels of that view (DatasetSamplers.py:10-41)."""

    def __init__(self, dataset, random: bool = True, img_sampler_cls=RandomImageSampler) -> None:
        self.mode = dataset.mode
        self.id_sampler = RandomSequentialSampler(len(dataset)) if random else SequentialSampler(len(dataset))
        self.img_samplers = [img_sampler_cls(v.camera.width * v.camera.height) for v in dataset] if img_sampler_cls else None

    def get(self, dataset, ray_batch_size: int | None = None) -> dict:
        if dataset.mode != self.mode:
            raise Framework.SamplerError(f'sampler initialised for mode "{self.mode}", dataset is in mode "{dataset.mode}"')
        sample_id = int(self.id_sampler.get(1).item())
        view = dataset[sample_id]
        image_sampler = ray_ids = ray_batch = None
        if self.img_samplers and ray_batch_size is not None:
            image_sampler = self.img_samplers[sample_id]
            ray_ids = image_sampler.get(ray_batch_size).to(Framework.config.GLOBAL.DEFAULT_DEVICE)
            collection = dataset.ray_collection[self.mode]
            if collection is not None:
                ray_batch = collection[sample_id][ray_ids]
            elif _takes_pixel_ids(view):
                ray_batch = view.get_rays(ray_ids)        # K0: only the sampled pixels
            else:
                ray_batch = view.get_rays()[ray_ids]      # a reference View: whole image, then index (DatasetSamplers.py:39)
        return {'sample_id': sample_id, 'view': view, 'image_sampler': image_sampler, 'ray_ids': ray_ids, 'ray_batch': ray_batch}


class RayPoolSampler:
    """Random rays from the pool of all training rays (DatasetSamplers.py:44-66)."""

    def __init__(self, dataset, img_sampler_cls=RandomImageSampler) -> None:
        self.mode = dataset.mode
        self.image_sampler = img_sampler_cls(dataset.get_total_ray_count())

    def get(self, dataset, ray_batch_size: int) -> dict:
        if dataset.mode != self.mode:
            raise Framework.SamplerError(f'sampler initialised for mode "{self.mode}", dataset is in mode "{dataset.mode}"')
        rays_all = dataset.get_all_rays()
        ray_ids = self.image_sampler.get(ray_batch_size).to(rays_all.device)
        return {'sample_id': None, 'view': None, 'image_sampler': self.image_sampler, 'ray_ids': ray_ids,
                'ray_batch': rays_all[ray_ids].to(device=Framework.config.GLOBAL.DEFAULT_DEVICE)}
