#include <GL/glew.h>
