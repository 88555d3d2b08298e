def dataclass(cls):
    return cls
