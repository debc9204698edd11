#include "../../cv_eigen_standin.hpp"
