def print(fmt, *a, **k):
    import builtins
    builtins.print(fmt.format(*a, **k))
