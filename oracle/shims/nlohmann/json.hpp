// Shim: forwards to the nlohmann/json 3.11.3 single header that ships inside the image
// (cudnn_frontend/thirdparty) -- the exact version the reference pins (CMakeLists.txt:39-43).
#pragma once
#include <cudnn_frontend/thirdparty/nlohmann/json.hpp>
