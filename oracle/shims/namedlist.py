"""ORACLE shim for namedlist.namedlist: mutable, iterable, picklable record type."""


def namedlist(name, fields):
    fields = list(fields)

    def __init__(self, *args):
        assert len(args) == len(fields)
        for f, a in zip(fields, args):
            setattr(self, f, a)

    def __iter__(self):
        return iter(getattr(self, f) for f in fields)

    def __reduce__(self):
        return (_rebuild, (name, tuple(fields), tuple(self)))

    cls = type(name, (object,), {"__slots__": tuple(fields), "__init__": __init__, "__iter__": __iter__,
                                 "__reduce__": __reduce__, "_fields": tuple(fields)})
    return cls


def _rebuild(name, fields, values):
    return namedlist(name, fields)(*values)
