"""reference: models/__init__.py:1-2 (`from models.pcn import PCN, VN_PCN`, `from models.dgcnn import DGCNN`)"""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)      # sub-modules not provided here fall through to the reference's models/

from models.pcn import PCN, VN_PCN  # noqa: E402,F401
from models.dgcnn import DGCNN  # noqa: E402,F401
