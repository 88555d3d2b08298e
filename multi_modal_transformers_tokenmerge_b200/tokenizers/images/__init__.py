from .image_tokenizer import ImageTokenizer, ResNetV2Block, encode_patch_position, image_to_patches_index  # noqa: F401
