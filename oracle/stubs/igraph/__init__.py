"""Empty stand-in: the reference imports igraph only for render() (evopf.py:21,391)."""
