from .lenet import Cnn, lenet_batches  # noqa: F401
