// bcd_b200 — stopwatch with the reference's public surface (superboubek/bcd include/bcd/core/Chronometer.h:20-45,
// behaviour of src/core/Chronometer.cpp:17-80), written from scratch and header-only: accumulating start/stop,
// seconds as float, and the "1234.5 ms ( 1s 234ms )" text the reference's tools print.
#ifndef BCD_B200_CHRONOMETER_H
#define BCD_B200_CHRONOMETER_H

#include <chrono>
#include <cmath>
#include <iostream>
#include <sstream>
#include <string>

namespace bcd
{

	class Chronometer
	{
	public:
		Chronometer() : m_isRunning(false), m_elapsedTime(0.f) {}

		void reset() { m_isRunning = false; m_elapsedTime = 0.f; }
		void start()
		{
			if(m_isRunning) return;
			m_isRunning = true;
			m_startTime = std::chrono::high_resolution_clock::now();
		}
		void stop()
		{
			if(!m_isRunning) return;
			m_elapsedTime = getElapsedTime();
			m_isRunning = false;
		}
		/// Seconds accumulated over all start/stop intervals (including the running one).
		float getElapsedTime() const
		{
			if(!m_isRunning) return m_elapsedTime;
			const std::chrono::duration<float> running = std::chrono::high_resolution_clock::now() - m_startTime;
			return m_elapsedTime + running.count();
		}
		/// "<ms> ms", followed by " ( [Hh ][Mmin ]Ss MSms )" from one second on.
		static std::string getStringFromTime(float i_timeInSeconds)
		{
			std::ostringstream oss;
			oss << i_timeInSeconds * 1000.f << " ms";
			if(i_timeInSeconds >= 1.f)
			{
				float rest = i_timeInSeconds;
				oss << " ( ";
				if(rest >= 3600.f) { const float h = std::floor(rest / 3600.f); oss << h << "h "; rest -= h * 3600.f; }
				if(rest >= 60.f) { const float m = std::floor(rest / 60.f); oss << m << "min "; rest -= m * 60.f; }
				const float sec = std::floor(rest);
				oss << sec << "s " << std::floor((rest - sec) * 1000.f) << "ms )";
			}
			return oss.str();
		}
		void printElapsedTime(std::ostream& o_stream = std::cout) { o_stream << getStringFromTime(getElapsedTime()); }

	private:
		bool m_isRunning;
		float m_elapsedTime;
		std::chrono::high_resolution_clock::time_point m_startTime;
	};

	inline std::ostream& operator<<(std::ostream& i_rOs, const Chronometer& i_rChrono)
	{
		return i_rOs << Chronometer::getStringFromTime(i_rChrono.getElapsedTime());
	}

} // namespace bcd

#endif // BCD_B200_CHRONOMETER_H
