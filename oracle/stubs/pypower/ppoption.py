def ppoption(*args, **kwargs):
    return dict(kwargs)
