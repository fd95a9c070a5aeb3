#pragma once
namespace indicators { inline void show_console_cursor(bool) {} }
