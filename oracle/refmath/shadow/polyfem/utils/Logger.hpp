// shadow of polyfem/utils/Logger.hpp for oracle/refmath: a logger that discards everything
#pragma once
#include <stdexcept>
#include <string>
namespace polyfem
{
	struct NullLogger
	{
		template <typename... A>
		void trace(A &&...) {}
		template <typename... A>
		void debug(A &&...) {}
		template <typename... A>
		void info(A &&...) {}
		template <typename... A>
		void warn(A &&...) {}
		template <typename... A>
		void error(A &&...) {}
	};
	inline NullLogger &logger()
	{
		static NullLogger l;
		return l;
	}
	template <typename... A>
	[[noreturn]] inline void log_and_throw_error(const std::string &msg, A &&...)
	{
		throw std::runtime_error(msg);
	}
} // namespace polyfem
