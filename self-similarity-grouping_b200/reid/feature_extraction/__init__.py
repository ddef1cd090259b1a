from .cnn import extract_cnn_feature  # noqa: F401

__all__ = ['extract_cnn_feature']
