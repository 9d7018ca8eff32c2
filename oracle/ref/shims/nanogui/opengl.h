/* include/matrix.h:3 of the reference only needs the GLfloat typedef. */
#pragma once
typedef float GLfloat;
