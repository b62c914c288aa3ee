// oracle/ref_stubs/brain-engine.h — TEST INFRASTRUCTURE ONLY.
// Shadows the reference's brain-engine.h (which pulls <Metal/Metal.hpp>) so that the reference's
// stimulus/functional-dataset.{h,cpp} compile verbatim on Linux (functional-dataset.h:6 includes it
// only to reach <functional>, <cstdint> and <vector>).
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <vector>
