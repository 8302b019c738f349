"""Visualisation helpers of the reference API (MTM/__init__.py:299-391).

Not on the compute path: thin pass-throughs to OpenCV's drawing primitives, kept
so scripts written against ``MTM.drawBoxesOnRGB/Gray`` keep working."""


def _overlay(canvas, listHit, thickness, box_colour, showLabel, label_colour, labelScale):
    import cv2
    for label, (x, y, w, h), _score in listHit:
        cv2.rectangle(canvas, (x, y), (x + w, y + h), color=box_colour, thickness=thickness)
        if showLabel:
            cv2.putText(canvas, text=label, org=(x, y), fontFace=cv2.FONT_HERSHEY_SIMPLEX,
                        fontScale=labelScale, color=label_colour, lineType=cv2.LINE_AA)
    return canvas


def drawBoxesOnRGB(image, listHit, boxThickness=2, boxColor=(255, 255, 0), showLabel=False,
                   labelColor=(255, 255, 0), labelScale=0.5):
    """Copy of ``image`` (converted to RGB when grayscale) with the hit boxes drawn."""
    import cv2
    canvas = cv2.cvtColor(image, cv2.COLOR_GRAY2RGB) if image.ndim == 2 else image.copy()
    return _overlay(canvas, listHit, boxThickness, boxColor, showLabel, labelColor, labelScale)


def drawBoxesOnGray(image, listHit, boxThickness=2, boxColor=255, showLabel=False, labelColor=255, labelScale=0.5):
    """Copy of ``image`` (converted to grayscale when RGB) with the hit boxes drawn."""
    import cv2
    canvas = cv2.cvtColor(image, cv2.COLOR_RGB2GRAY) if image.ndim == 3 else image.copy()
    return _overlay(canvas, listHit, boxThickness, boxColor, showLabel, labelColor, labelScale)
