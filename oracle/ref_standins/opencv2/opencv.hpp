#include "../cv_standin.hpp"
