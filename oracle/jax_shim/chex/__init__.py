"""chex stand-in: the one assertion tokenizers/images/image_tokenizer.py uses."""


def assert_equal(a, b):
    assert a == b, (a, b)
