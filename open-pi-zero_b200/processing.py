"""`VLAProcessor` of the reference (`src/model/vla/processing.py:63-136`): builds the PaliGemma prompt
(`<image>` x num_image_tokens + <bos> + text + "\\n"), runs the caller's HuggingFace-style tokenizer with the reference's
arguments, and prepares the pixels.  Same constructor, same `__call__(text, images, truncation)`, same output keys.

B200-native difference: `keep_uint8=True` returns the camera frames as they are (uint8) -- the rescale / normalise of
`process_images` (processing.py:47-60) then runs inside the patch-gather kernel (`pz_set_pixel_format(PZ_PIXELS_U8)`,
csrc/elementwise.cu `im2col_u8_kernel`), a quarter of the host-to-device bytes and no host arithmetic.  The tokenizer itself
is the third-party `transformers` object the caller loads (the PaliGemma tokenizer files are not part of either repo)."""
from __future__ import annotations

from typing import List

import torch

IMAGENET_STANDARD_MEAN = torch.tensor([0.5, 0.5, 0.5])
IMAGENET_STANDARD_STD = torch.tensor([0.5, 0.5, 0.5])


def add_image_tokens_to_prompt(prefix_prompt, bos_token, image_seq_len, image_token):
    """processing.py:9-23."""
    return f"{image_token * image_seq_len}{bos_token}{prefix_prompt}\n"


def process_images(images: torch.Tensor, rescale_factor: float, image_mean: torch.Tensor, image_std: torch.Tensor) -> torch.Tensor:
    """processing.py:26-60: rescale to [0, 1], then (x - mean) / std per channel."""
    assert images.ndim == 4, f"Expected 4D tensor, got {images.ndim}D tensor."
    assert images.shape[1] == 3, f"Expected 3 channels at axis 1, got {images.shape[1]} channels."
    images = images * rescale_factor
    return (images - image_mean[None, :, None, None]) / image_std[None, :, None, None]


class VLAProcessor:
    IMAGE_TOKEN = "<image>"

    def __init__(self, tokenizer, num_image_tokens: int, max_seq_len: int, tokenizer_padding: str = "max_length",
                 keep_uint8: bool = False):
        self.image_seq_length = num_image_tokens
        self.max_seq_len = max_seq_len
        self.tokenizer_padding = tokenizer_padding
        self.keep_uint8 = keep_uint8
        tokenizer.add_special_tokens({"additional_special_tokens": [self.IMAGE_TOKEN]})
        extra = [f"<loc{i:04d}>" for i in range(1024)]      # object detection (bounding boxes)
        extra += [f"<seg{i:03d}>" for i in range(128)]      # object segmentation
        tokenizer.add_tokens(extra)
        self.image_token_id = tokenizer.convert_tokens_to_ids(self.IMAGE_TOKEN)
        tokenizer.add_bos_token = False                      # BOS / EOS are added by the prompt builder
        tokenizer.add_eos_token = False
        self.tokenizer = tokenizer

    def __call__(self, text: List[str], images: torch.Tensor, truncation: bool = True) -> dict:
        assert len(images) == len(text), f"Received {len(images)} images for {len(text)} prompts."
        assert images.dtype == torch.uint8, f"Expected uint8 tensor for images, got {images.dtype}."
        if self.keep_uint8:
            pixel_values = images          # normalised inside the patch-gather kernel
        else:
            pixel_values = process_images(images, rescale_factor=1 / 255.0, image_mean=IMAGENET_STANDARD_MEAN,
                                          image_std=IMAGENET_STANDARD_STD)
        input_strings = [add_image_tokens_to_prompt(prefix_prompt=prompt, bos_token=self.tokenizer.bos_token,
                                                    image_seq_len=self.image_seq_length, image_token=self.IMAGE_TOKEN)
                         for prompt in text]
        inputs = self.tokenizer(input_strings, return_tensors="pt", max_length=self.max_seq_len,
                                padding=self.tokenizer_padding, truncation=truncation)
        return {"pixel_values": pixel_values, **inputs}
