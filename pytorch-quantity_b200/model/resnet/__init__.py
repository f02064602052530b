from .resnet_fabu import resnet18_fabu, resnet50_fabu, FabuResNet  # noqa: F401
