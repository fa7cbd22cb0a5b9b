// stub: RVI/parameter/parameters.h includes OpenCV for types the factor sources never use
