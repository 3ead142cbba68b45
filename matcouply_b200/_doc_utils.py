"""Source compatibility with custom penalties written for the reference (``matcouply/_doc_utils.py``): user subclasses
decorate their methods with ``@copy_ancestor_docstring`` (``examples/plot_custom_penalty.py:213-231``).  Here the
decorator copies the docstring of the first ancestor method it can find at class-creation time when the class uses
``InheritableDocstrings``, and is a plain no-op otherwise — it never raises, so the import line of such a penalty only
needs the package name changed."""
from abc import ABCMeta

from .penalties import copy_ancestor_docstring  # noqa: F401  (no-op decorator)

__all__ = ["copy_ancestor_docstring", "InheritableDocstrings"]


class InheritableDocstrings(ABCMeta):
    """Metaclass: methods without a docstring inherit the one of the same-named method of the nearest ancestor
    (what the reference's metaclass + decorator pair achieves, ``_doc_utils.py:74-102``)."""

    def __new__(mcls, name, bases, classdict):
        cls = super().__new__(mcls, name, bases, classdict)
        for attr, fn in classdict.items():
            if callable(fn) and getattr(fn, "__doc__", None) is None:
                for base in cls.__mro__[1:]:
                    doc = getattr(getattr(base, attr, None), "__doc__", None)
                    if doc:
                        try:
                            fn.__doc__ = doc
                        except (AttributeError, TypeError):
                            pass
                        break
        return cls
