// source.hpp — the pull-style byte source every input layer of the compiled host implements.
#pragma once
#include <cstddef>
#include <cstdint>

namespace host {

// read() fills up to n bytes and returns how many; 0 means end of input.  Errors are exceptions.
struct Source {
    virtual ~Source() {}
    virtual size_t read(uint8_t *dst, size_t n) = 0;
};

}  // namespace host
