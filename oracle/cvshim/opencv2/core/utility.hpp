/* cvshim: everything the reference needs lives in opencv2/core.hpp (test infrastructure, see there). */
#include <opencv2/core.hpp>
