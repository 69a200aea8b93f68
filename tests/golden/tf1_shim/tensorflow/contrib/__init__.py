"""tf.contrib stand-in (see ../__init__.py): only contrib.layers is used by the reference's hot path."""
from . import layers  # noqa: F401
