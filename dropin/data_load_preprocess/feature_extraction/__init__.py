"""Drop-in sub-package (see ``data_load_preprocess/__init__.py``)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
