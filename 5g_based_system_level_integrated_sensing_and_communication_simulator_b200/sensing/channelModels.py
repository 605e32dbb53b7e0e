"""``+sensing/+channelModels`` mirror."""
from ._echo import basicRadarChannel  # noqa: F401
