"""dream.models -> dream_b200.models (DreamHourglass, DreamHourglassMultiStage, ResnetSimple; dream/models.py)."""
from dream_b200.models import DreamHourglass, DreamHourglassMultiStage, ResnetSimple   # noqa: F401
