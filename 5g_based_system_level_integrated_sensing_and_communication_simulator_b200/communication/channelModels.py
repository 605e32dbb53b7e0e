"""``+communication/+channelModels`` mirror."""


def updateCDLModels(simuParams):
    """``delayProfile = communication.channelModels.updateCDLModels(simuParams)`` (updateCDLModels.m:7-15):
    CDL-D for LoS links, CDL-A otherwise (per UE)."""
    return ["CDL-D" if int(los) == 1 else "CDL-A" for los in simuParams["ueLoSConditions"]]
